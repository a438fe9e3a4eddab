// ahmc_tree_logic.cuh -- the SCALAR half of the iterative NUTS transition as a per-chain state machine.
//
// STATUS: prototype for the tile-centric NUTS kernel planned next ("K3D").  It is NOT part of
// libahmc_b200.so and no product kernel includes it yet.  It compiles for host and device; its logic is exercised on
// the CPU by tests/test_tree_logic.py (a host harness supplies the vector work) against the recursive C oracle for
// both trajectory samplers and all three termination criteria.  Nothing here has run on a GPU.
//
// Why a state machine: in K3 every lane of the warp that owns a chain repeats the chain's scalar bookkeeping (energy
// error, exp / logaddexp, random draw, multinomial or slice decision, binary-counter state, termination flags).  In the
// tile design ONE warp does that bookkeeping with lane = chain while the other warps do vector work on [D x 32] tiles;
// the two halves talk through the small set of per-chain actions and reduced scalars defined here.
//
// Semantics = advancedhmc.jl_b200/csrc/ahmc_nuts_kernel.cuh = src/trajectory.jl:626-742:
//   leaf weights H0 - H' (:174-176) or slice counts (:164-166); one variate per internal combine in post-order
//   (:191-195 / :178-183) and one for the top-level mh_accept only if the subtree did not terminate (:708-713); a
//   terminated first half is returned without building its sibling (:652) -- it "floats" up through levels whose bit is
//   0; divergence iff !(-H0 < delta_max - H') (:503-507) resp. !(lu < delta_max - H') (:500-502).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define AHMC_HD __host__ __device__
#else
#define AHMC_HD
#endif

namespace ahmc {
namespace tree {

enum Action : int {
    ACT_LEAF = 0,      // take one leapfrog step in direction v, then call after_leaf()
    ACT_COMBINE = 1,   // vector half: rho_cur += pending[k].rho and the U-turn dots of the merged node; then combine()
    ACT_STORE = 2,     // vector half: park the current node as pending[k]; then stored()
    ACT_COMPLETE = 3,  // vector half: top-level merge (edge, rho_tree, dots); then complete()
    ACT_DONE = 4       // the transition is finished: zcand / statistics are final
};

enum Sampler : int { MULTINOMIAL = 0, SLICE = 1 };
enum Criterion : int { GENERALISED = 0, CLASSIC = 1, STRICT = 2 };

AHMC_HD inline double jl_min0(double x) { return (x != x) ? x : (x < 0.0 ? x : 0.0); }  // min(0, x) like Julia
AHMC_HD inline double logaddexp(double a, double b) {
    const double delta = (a == b) ? 0.0 : fabs(a - b);
    const double mx = (a != a || b != b) ? NAN : (a > b ? a : b);
    return mx + log1p(exp(-delta));
}
AHMC_HD inline double maxabs(double a, double b) { return fabs(a) > fabs(b) ? a : b; }  // :526

// U-turn inputs the vector half reduces for a merge of (first-built F, second-built S) or for the top level
// (tree T, new subtree S); which entries are needed depends on the criterion.
struct Dots {
    double g1, g2;          // Generalised: rho_merged . M^-1 r_left-or-first, rho_merged . M^-1 r_leaf
    double c1, c2;          // Classic:     q . M^-1 r_first(or far edge),     q . M^-1 r_leaf,  q = theta_left - theta_right
    double a1, a2, b1, b2;  // Strict extras (DESIGN / kernel comments: checks A and B, resp. X and Y at the top level)
};

AHMC_HD inline bool uturn(int criterion, const Dots& d) {
    if (criterion == CLASSIC) return (d.c1 >= 0.0) || (d.c2 >= 0.0);                   // :551-557
    bool s = (d.g1 <= 0.0) || (d.g2 <= 0.0);                                          // :566-570, :615-617
    if (criterion == STRICT) s = s || (d.a1 <= 0.0) || (d.a2 <= 0.0) || (d.b1 <= 0.0) || (d.b2 <= 0.0);  // :579-613
    return s;
}

template <int MAXD>
struct Chain {
    // ---- configuration
    int sampler, criterion, max_depth;
    double delta_max;
    // ---- transition state (:682-688)
    double H0, lu, lw_tree, sa_tree, dh_tree;
    int na_tree, j;
    bool term_dyn, term_num;
    // ---- subtree under construction
    int v, jsub, i, k;
    // ---- current node (the leaf just built, then whatever it has been merged into)
    double lw_c, sa_c, na_c, dh_c;
    bool tnum_c, tdyn_c;
    int cand_cur;   // -1: the candidate is the current leaf; k >= 0: the candidate parked in level k
    int cand_out;   // where the transition's candidate lives after an accept: -2 unchanged, -1 leaf, k level (one shot)
    // ---- pending levels: scalars of the first half-subtree waiting at level k
    double LW[MAXD], SA[MAXD], NA[MAXD], DH[MAXD];
    int n_var, n_dir;  // variates / direction bits consumed so far

    // begin a transition at z0 with energy H0 = -(lp + lk); `e0` = the first variate (randexp) is consumed by SliceTS only
    AHMC_HD void begin(int sampler_, int criterion_, int max_depth_, double delta_max_, double neg_energy0) {
        sampler = sampler_;
        criterion = criterion_;
        max_depth = max_depth_;
        delta_max = delta_max_;
        H0 = -neg_energy0;
        lu = 0.0;
        lw_tree = 0.0;
        sa_tree = dh_tree = 0.0;
        na_tree = j = 0;
        term_dyn = term_num = false;
        n_var = n_dir = 0;
        cand_out = -2;
    }
    AHMC_HD bool needs_slice_variate() const { return sampler == SLICE && n_var == 0; }
    AHMC_HD void slice_init(double randexp) {  // SliceTS(rng, z0) (:144-145)
        lu = -H0 - randexp;
        lw_tree = 1.0;
        ++n_var;
    }
    AHMC_HD bool finished() const { return term_dyn || term_num || !(j < max_depth); }

    // (A) start a doubling with direction bit `left` (rand(rng, Bool), :693); the vector half then loads the edge
    AHMC_HD void start_doubling(bool left) {
        v = left ? -1 : 1;
        jsub = j;
        i = 0;
        ++n_dir;
    }

    // (B) after the leaf z' = step(z, v): neg_energy(z') -> the node's scalars; returns the first merge action
    AHMC_HD int after_leaf(double nE) {
        const double H1 = -nE, dH = H1 - H0;
        if (sampler == SLICE) {
            lw_c = (lu <= nE) ? 1.0 : 0.0;            // :164-166
            tnum_c = !(lu < delta_max + -H1);         // :500-502
        } else {
            lw_c = H0 + nE;                           // :174-176
            tnum_c = !(-H0 < delta_max + -H1);        // :503-507
        }
        sa_c = exp(jl_min0(-dH));
        na_c = 1.0;
        dh_c = dH;
        tdyn_c = false;
        cand_cur = -1;
        k = 0;
        return next_action();
    }

    // (C) what the binary counter asks for at level k (floats are taken here: they need no vector work)
    AHMC_HD int next_action() {
        while (true) {
            if (k == jsub) return ACT_COMPLETE;
            if ((i >> k) & 1) return ACT_COMBINE;
            if (tnum_c || tdyn_c) {  // terminated first half: returned as is (:652)
                ++k;
                continue;
            }
            return ACT_STORE;
        }
    }

    // combine with the half pending at level k; `var` = randexp (Multinomial) or rand (Slice).  Returns true when the
    // merged node's candidate is the PENDING half's (the vector half then remembers cand_cur = k).
    AHMC_HD bool combine(const Dots& d, double var) {
        const double lw_p = LW[k], sa_p = SA[k], na_p = NA[k], dh_p = DH[k];
        ++n_var;
        bool take_pending;
        if (sampler == SLICE) {  // :178-183
            const double n = lw_p + lw_c;
            take_pending = n * var < lw_p;
            lw_c = n;
        } else {  // :191-195
            const double lw = logaddexp(lw_p, lw_c);
            take_pending = lw < lw_p + var;
            lw_c = lw;
        }
        if (take_pending) cand_cur = k;
        sa_c = (v > 0) ? sa_p + sa_c : sa_c + sa_p;  // treeleft + treeright (:538)
        na_c += na_p;
        dh_c = (v > 0) ? maxabs(dh_p, dh_c) : maxabs(dh_c, dh_p);
        tdyn_c = tdyn_c || uturn(criterion, d);
        ++k;
        return take_pending;
    }

    // the vector half has parked the node at level k: remember its scalars, move on to the next leaf
    AHMC_HD void stored() {
        LW[k] = lw_c;
        SA[k] = sa_c;
        NA[k] = na_c;
        DH[k] = dh_c;
        ++i;
    }

    // (D) subtree complete: `var` is consumed only if the subtree did not terminate (:708-713).
    // Returns the next action (ACT_LEAF after start_doubling() by the caller, or ACT_DONE); sets cand_out.
    AHMC_HD bool subtree_terminated() const { return tnum_c || tdyn_c; }
    AHMC_HD void complete(const Dots& d, double var) {
        cand_out = -2;
        if (!subtree_terminated()) {
            j = j + 1;
            ++n_var;
            const bool accept = (sampler == SLICE) ? (lw_tree * var < lw_c) : (lw_tree < lw_c + var);  // :202, :204-206
            if (accept) cand_out = cand_cur;
        }
        sa_tree = (v < 0) ? sa_c + sa_tree : sa_tree + sa_c;
        na_tree += (int)na_c;
        dh_tree = (v < 0) ? maxabs(dh_c, dh_tree) : maxabs(dh_tree, dh_c);
        lw_tree = (sampler == SLICE) ? lw_tree + lw_c : logaddexp(lw_tree, lw_c);  // :185-189, :197-200
        term_dyn = term_dyn || tdyn_c || uturn(criterion, d);                       // :719-722
        term_num = term_num || tnum_c;
    }
};

}  // namespace tree
}  // namespace ahmc
