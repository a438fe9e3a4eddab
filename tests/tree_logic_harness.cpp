// tree_logic_harness.cpp -- CPU harness for advancedhmc.jl_b200/csrc/experimental/ahmc_tree_logic.cuh (test infrastructure).
// The state machine owns the scalars and the control flow; this file plays the "vector half" of the planned tile kernel
// for a diagonal-Gaussian target with a diagonal metric: leapfrog steps, the per-level vector workspace (rho, r_first,
// r_last, theta_first, candidate), U-turn dots.  tests/test_tree_logic.py compares its transitions with the recursive C
// oracle on shared tapes.  Built by the test with g++ (no CUDA involved).
#include <cstdint>
#include <cstring>
#include <vector>

#include "ahmc_tree_logic.cuh"

using namespace ahmc::tree;
constexpr int kMaxD = 12;

namespace {
struct Vec {
    std::vector<double> th, r, g;  // g = -grad log pi
    double lp = 0, lk = 0;
};
struct Sys {
    int D;
    const double *mu, *sd, *Minv;
    double eps;
    void grad(Vec& z) const {
        double s = 0;
        for (int d = 0; d < D; ++d) {
            const double diff = z.th[d] - mu[d], w = 1.0 / (sd[d] * sd[d]);
            z.g[d] = diff * w;
            s += diff * diff * w;
        }
        z.lp = -0.5 * s;
    }
    void kinetic(Vec& z) const {
        double s = 0;
        for (int d = 0; d < D; ++d) s += z.r[d] * Minv[d] * z.r[d];
        z.lk = -0.5 * s;
    }
    void step(Vec& z, int v) const {  // integrator.jl:216-265, n_steps = v
        const double e = v > 0 ? eps : -eps;
        for (int d = 0; d < D; ++d) z.r[d] -= e / 2 * z.g[d];
        for (int d = 0; d < D; ++d) z.th[d] += e * Minv[d] * z.r[d];
        grad(z);
        for (int d = 0; d < D; ++d) z.r[d] -= e / 2 * z.g[d];
        kinetic(z);
    }
    double dotM(const std::vector<double>& a, const std::vector<double>& r) const {  // a . M^-1 r
        double s = 0;
        for (int d = 0; d < D; ++d) s += a[d] * Minv[d] * r[d];
        return s;
    }
};
struct LevelVecs {
    std::vector<double> rho, rfirst, rlast, thfirst;
    Vec cand;
};
}  // namespace

extern "C" int tree_logic_nuts(int D, int64_t N, const double* mu, const double* sd, const double* Minv, double eps,
                               int sampler, int criterion, int max_depth, double delta_max, const double* theta0,
                               const double* r0, const uint8_t* dirs, int64_t dir_stride, const double* var,
                               int64_t var_stride, double* theta_out, double* r_out, int32_t* n_steps, int32_t* depth,
                               int32_t* numerical, double* acc_rate, double* dh_max, int32_t* vars_used) {
    if (max_depth > kMaxD) return -1;
    Sys S{D, mu, sd, Minv, eps};
    for (int64_t c = 0; c < N; ++c) {
        const uint8_t* dt = dirs + dir_stride * c;
        const double* vt = var + var_stride * c;
        Vec z0;
        z0.th.assign(theta0 + D * c, theta0 + D * (c + 1));
        z0.r.assign(r0 + D * c, r0 + D * (c + 1));
        z0.g.resize(D);
        S.grad(z0);
        S.kinetic(z0);
        Chain<kMaxD> T;
        T.begin(sampler, criterion, max_depth, delta_max, z0.lp + z0.lk);
        if (T.needs_slice_variate()) T.slice_init(vt[T.n_var]);
        Vec LEFT = z0, RIGHT = z0, zcand = z0, cur;
        std::vector<double> rho_tree = z0.r;
        std::vector<LevelVecs> lv(max_depth > 0 ? max_depth : 1);
        // the current node's vectors
        std::vector<double> rho_cur, rfirst_cur, thfirst_cur;
        while (!T.finished()) {
            T.start_doubling(dt[T.n_dir] != 0);
            cur = T.v < 0 ? LEFT : RIGHT;
            bool complete = false;
            while (!complete) {
                S.step(cur, T.v);  // ACT_LEAF
                int act = T.after_leaf(cur.lp + cur.lk);
                rho_cur = cur.r;
                rfirst_cur = cur.r;
                thfirst_cur = cur.th;
                while (true) {
                    if (act == ACT_COMBINE) {
                        LevelVecs& F = lv[T.k];
                        Dots d{};
                        if (criterion == STRICT) {  // checks A and B on (first-built F, second-built S = current)
                            std::vector<double> ra(D), rb(D);
                            for (int q = 0; q < D; ++q) {
                                ra[q] = F.rho[q] + rfirst_cur[q];
                                rb[q] = rho_cur[q] + F.rlast[q];
                            }
                            d.a1 = S.dotM(ra, F.rfirst);
                            d.a2 = S.dotM(ra, rfirst_cur);
                            d.b1 = S.dotM(rb, cur.r);
                            d.b2 = S.dotM(rb, F.rlast);
                        }
                        for (int q = 0; q < D; ++q) rho_cur[q] += F.rho[q];
                        if (criterion == CLASSIC) {
                            std::vector<double> qv(D);
                            for (int q = 0; q < D; ++q) qv[q] = T.v > 0 ? F.thfirst[q] - cur.th[q] : cur.th[q] - F.thfirst[q];
                            d.c1 = S.dotM(qv, F.rfirst);
                            d.c2 = S.dotM(qv, cur.r);
                        } else {
                            d.g1 = S.dotM(rho_cur, F.rfirst);
                            d.g2 = S.dotM(rho_cur, cur.r);
                        }
                        const int kk = T.k;
                        T.combine(d, vt[T.n_var]);
                        rfirst_cur = lv[kk].rfirst;   // the merged node's first-built leaf is F's
                        thfirst_cur = lv[kk].thfirst;
                        act = T.next_action();
                    } else if (act == ACT_STORE) {
                        LevelVecs& L = lv[T.k];
                        L.rho = rho_cur;
                        L.rfirst = rfirst_cur;
                        L.thfirst = thfirst_cur;
                        L.rlast = cur.r;
                        if (T.cand_cur < 0) L.cand = cur;
                        else if (T.cand_cur != T.k) L.cand = lv[T.cand_cur].cand;
                        T.stored();
                        break;  // next leaf
                    } else {  // ACT_COMPLETE
                        const Vec& near = T.v < 0 ? LEFT : RIGHT;
                        const Vec& far = T.v < 0 ? RIGHT : LEFT;
                        Dots d{};
                        if (criterion == STRICT) {  // checks X and Y on (tree T, new subtree S)
                            std::vector<double> rx(D), ry(D);
                            for (int q = 0; q < D; ++q) {
                                rx[q] = rho_tree[q] + rfirst_cur[q];
                                ry[q] = near.r[q] + rho_cur[q];
                            }
                            d.a1 = S.dotM(rx, far.r);
                            d.a2 = S.dotM(rx, rfirst_cur);
                            d.b1 = S.dotM(ry, near.r);
                            d.b2 = S.dotM(ry, cur.r);
                        }
                        for (int q = 0; q < D; ++q) rho_tree[q] += rho_cur[q];
                        if (criterion == CLASSIC) {
                            std::vector<double> qv(D);
                            for (int q = 0; q < D; ++q) qv[q] = T.v > 0 ? far.th[q] - cur.th[q] : cur.th[q] - far.th[q];
                            d.c1 = S.dotM(qv, far.r);
                            d.c2 = S.dotM(qv, cur.r);
                        } else {
                            d.g1 = S.dotM(rho_tree, far.r);
                            d.g2 = S.dotM(rho_tree, cur.r);
                        }
                        const double v_top = T.subtree_terminated() ? 0.0 : vt[T.n_var];
                        T.complete(d, v_top);
                        if (T.cand_out == -1) zcand = cur;
                        else if (T.cand_out >= 0) zcand = lv[T.cand_out].cand;
                        if (T.v < 0) LEFT = cur;
                        else RIGHT = cur;
                        complete = true;
                        break;
                    }
                }
            }
        }
        std::memcpy(theta_out + D * c, zcand.th.data(), sizeof(double) * D);
        std::memcpy(r_out + D * c, zcand.r.data(), sizeof(double) * D);
        n_steps[c] = T.na_tree;
        depth[c] = T.j;
        numerical[c] = T.term_num ? 1 : 0;
        acc_rate[c] = T.sa_tree / (double)T.na_tree;
        dh_max[c] = T.dh_tree;
        vars_used[c] = T.n_var;
    }
    return 0;
}
