import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, ahmc_b200 as A
sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
from run_configs import correlated_gaussian, SEED
dev = "cuda:0"
D, N = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 8192
Sigma, P = correlated_gaussian(D, SEED + 5)
h = A.Hamiltonian(A.DenseEuclideanMetric(Sigma), A.DenseGaussian(np.zeros(D), P))
kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.5), A.GeneralisedNoUTurn()))
th = torch.as_tensor(np.random.default_rng(0).normal(size=(N, D)), device=dev)
z = A.phasepoint(h, th, torch.zeros_like(th))
for _ in range(3):
    tr = A.transition(A.PhiloxRNG(1), h, kern, z)
print("steps", tr.stat["n_steps"].double().mean().item())
