"""Times the per-step (exact) trajectory path on a few targets with the library named by AHMC_B200_LIB (A/B runs of two
builds): headline shape with AHMC_FLAG_EXACT_CHECKS, tempered leapfrog, Neal's funnel, and a run-time compiled user target.
Prints one JSON line; CUDA events on the context's stream, 30 launches after 5 warm-ups."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ahmc_b200 as A

USER_DIAG = r'''
#define AHMC_USER_COORDWISE
__device__ double ahmc_user_coord(int d, double x, const double* p, double* gd) {
    const double diff = x - p[2 * d], g = diff * p[2 * d + 1];
    *gd = -g;
    return -0.5 * diff * g;
}
'''
USER_LOGISTIC = r'''
#define AHMC_USER_COORDWISE
__device__ double ahmc_user_coord(int d, double x, const double* p, double* gd) {   // product of logistic densities
    const double e = exp(-x);
    *gd = -1.0 + 2.0 * e / (1.0 + e);
    return -x - 2.0 * log1p(e);
}
'''


def main():
    dev = torch.device("cuda:0")
    N, D, Ls = 4096, 128, 32
    ctx = A.get_context(0)
    stream = ctx.torch_stream()
    rng = np.random.default_rng(0)
    s = np.exp(np.linspace(np.log(0.1), np.log(10), D))
    out = {"lib": os.environ.get("AHMC_B200_LIB", "in-tree")}

    def timed(fn, reps=30):
        with torch.cuda.stream(stream):
            for _ in range(5):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3  # us

    def pp(h, Dd, scale):
        g = torch.Generator(device=dev).manual_seed(1)
        return A.phasepoint(h, scale * torch.randn((N, Dd), generator=g, dtype=torch.float64, device=dev),
                            torch.randn((N, Dd), generator=g, dtype=torch.float64, device=dev))

    h = A.Hamiltonian(A.DiagEuclideanMetric(s * s), A.DiagGaussian(np.zeros(D), s))
    z = pp(h, D, 1.0)
    out["exact_us"] = timed(A.StepPlan(A.Leapfrog(0.1), h, z, Ls, flags=A.FLAG_ASYNC | A.FLAG_EXACT_CHECKS))
    out["fast_us"] = timed(A.StepPlan(A.Leapfrog(0.1), h, z, Ls, flags=A.FLAG_ASYNC))
    out["tempered_us"] = timed(A.StepPlan(A.TemperedLeapfrog(0.1, 1.05), h, z, Ls, flags=A.FLAG_ASYNC))
    hf = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(100)), A.Funnel(100))
    out["funnel_us"] = timed(A.StepPlan(A.Leapfrog(0.05), hf, pp(hf, 100, 0.5), Ls, flags=A.FLAG_ASYNC))
    hf128 = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(128)), A.Funnel(128))
    out["funnel128_us"] = timed(A.StepPlan(A.Leapfrog(0.05), hf128, pp(hf128, 128, 0.5), Ls, flags=A.FLAG_ASYNC))
    hu = A.Hamiltonian(A.DiagEuclideanMetric(s * s), A.UserTarget(D, USER_DIAG, params=np.stack([np.zeros(D), 1.0 / s ** 2], axis=1)))
    out["user_diag_us"] = timed(A.StepPlan(A.Leapfrog(0.1), hu, pp(hu, D, 1.0), Ls, flags=A.FLAG_ASYNC))
    hl = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(D)), A.UserTarget(D, USER_LOGISTIC))
    out["user_logistic_us"] = timed(A.StepPlan(A.Leapfrog(0.1), hl, pp(hl, D, 1.0), Ls, flags=A.FLAG_ASYNC))
    # static HMC transitions (K2) on the funnel: 20 per launch
    kern = A.HMCKernel(A.Trajectory(A.EndPointTS, A.Leapfrog(0.05), A.FixedNSteps(Ls)))
    prn = A.PhiloxRNG(5)
    zf = pp(hf, 100, 0.5)
    out["funnel_hmc_us_per_transition"] = timed(lambda: A.sample_transitions(prn, hf, kern, zf, 20, keep_draws=False, flags=A.FLAG_ASYNC), reps=5) / 20
    print(json.dumps(out))


if __name__ == "__main__":
    main()
