"""Pose sampler of the bilevel loop, batched: all K poses of a step in one set of tensor ops.

Mirrors optimization/utils/load_LINEMOD_noscale.py (LL) and optimization/utils/gumble.py (GU):

  pose_spherical(theta, phi, radius)                 LL:63-72  (differentiable in phi)   -> [K, 4, 4]
  sample_pose_nograd(categorical_prob, num_K, T)     LL:250-301                          -> poses, sample_log
  sample_pose(categorical_prob, num_K, T, log)       LL:202-247                          -> poses (graph-attached to psi)

The reference builds each pose with a Python loop of 4x4 matmuls on the host and stacks them; here the K Gumbel-softmax
samples (GU:58-63), the uniform jitter (LL:233-236) and the closed form of  P @ R_theta @ R_phi @ T(radius)  are evaluated
for all K at once on whatever device `categorical_prob` lives on, so the psi -> c2w -> rays -> render graph stays on the
GPU and `render_path_grad` pulls dL/dpsi through it with one autograd call per image.

Closed form (c = cos phi, s = sin phi, ct = cos theta, st = sin theta, r = radius); same association as the reference's
matmul chain, so the entries round identically:

    R_phi @ T      = [[1, 0, 0, 0], [0, c, -s, -s r], [0, s, c, c r], [0, 0, 0, 1]]
    R_theta @ ...  = [[ct, -st s, -st c, -st (c r)], [0, c, -s, -s r], [st, ct s, ct c, ct (c r)], [0, 0, 0, 1]]
    P @ ...        = rows (-row0, row2, row1, row3)
"""
import numpy as np
import torch

BIN_WIDTH = 45.0
RADIUS = 1.01                      # LL:244 / LL:293
THETA_RANGE = (85.0, 95.0)         # LL:292


def bin_centres(n_cats=8, device=None):
    """LL:215: centres of the phi bins, [0, 45, ...] + 22.5 degrees."""
    return torch.arange(n_cats, dtype=torch.float32, device=device) * BIN_WIDTH + BIN_WIDTH / 2


def pose_spherical(theta, phi, radius):
    """LL:63-72 for a batch: theta [K] (degrees, constant), phi [K] (degrees, may carry grad) -> c2w [K, 4, 4]."""
    phi = torch.as_tensor(phi, dtype=torch.float32)
    theta = torch.as_tensor(theta, dtype=torch.float32, device=phi.device)
    phi, theta = phi.reshape(-1), theta.reshape(-1)
    r = torch.as_tensor(radius, dtype=torch.float32, device=phi.device)
    ph = phi / 180. * np.pi
    th = theta / 180. * np.pi
    c, s = torch.cos(ph), torch.sin(ph)
    ct, st = torch.cos(th), torch.sin(th)
    zero, one = torch.zeros_like(c), torch.ones_like(c)
    cr, sr = c * r, -s * r
    row0 = torch.stack([ct, -st * s, -st * c, -st * cr], -1)
    row1 = torch.stack([zero, c, -s, sr], -1)
    row2 = torch.stack([st, ct * s, ct * c, ct * cr], -1)
    row3 = torch.stack([zero, zero, zero, one], -1)
    return torch.stack([-row0, row2, row1, row3], -2)


def gumbel_softmax_angles(logits, degrees, gumbel_noises, temperature):
    """GU:58-63 for K noise vectors at once: softmax((logits + g_k) / T) . degrees -> [K]."""
    y = torch.softmax((logits[None, :] + gumbel_noises) / temperature, dim=-1)
    return (y * degrees[None, :]).sum(-1)


def sample_pose(categorical_prob, num_K, gumble_T, sample_log):
    """LL:202-247: replay the logged noise differentiably; poses [num_K, 4, 4] on categorical_prob's device."""
    dev = categorical_prob.device
    degrees = bin_centres(len(categorical_prob), dev)
    logits = torch.log(categorical_prob)
    g = torch.as_tensor(np.asarray(sample_log['gumbel_noises'][:num_K]), dtype=torch.float32, device=dev)
    u = torch.as_tensor(np.asarray(sample_log['uniform_noises'][:num_K]), dtype=torch.float32, device=dev)
    thetas = torch.as_tensor(np.asarray(sample_log['thetas'][:num_K]), dtype=torch.float32, device=dev)
    phi = gumbel_softmax_angles(logits, degrees, g, gumble_T) - BIN_WIDTH / 2 + BIN_WIDTH * u
    return pose_spherical(thetas, phi - 180, RADIUS)


def sample_pose_nograd(categorical_prob, num_K, gumble_T, seed=None, device=None):
    """LL:250-301.  Draw order follows the reference (K Gumbel vectors, then K uniforms, then K thetas) so a given numpy
    seed produces the same sample_log; the reference seeds with the wall-clock second (LL:272), `seed=None` does the same."""
    probs = np.asarray(categorical_prob.detach().cpu() if torch.is_tensor(categorical_prob) else categorical_prob,
                       dtype=np.float64)
    if seed is None:
        from datetime import datetime
        seed = datetime.now().second
    rng = np.random.RandomState(seed)
    n = len(probs)
    gumbel = np.stack([rng.gumbel(size=n) for _ in range(num_K)], 0)
    uniform = np.array([rng.uniform(0, 1) for _ in range(num_K)])
    thetas = np.array([rng.uniform(*THETA_RANGE) for _ in range(num_K)])
    degrees = np.arange(n) * BIN_WIDTH + BIN_WIDTH / 2
    z = (np.log(probs)[None, :] + gumbel) / gumble_T
    y = np.exp(z) / np.exp(z).sum(-1, keepdims=True)                   # GU:42-43 (no max subtraction, as there)
    phi = (y * degrees).sum(-1) - BIN_WIDTH / 2 + BIN_WIDTH * uniform
    poses = pose_spherical(torch.as_tensor(thetas, dtype=torch.float32, device=device),
                           torch.as_tensor(phi - 180, dtype=torch.float32, device=device), RADIUS).detach()
    log = {'gumbel_noises': gumbel.tolist(), 'uniform_noises': uniform.tolist(), 'thetas': thetas.tolist()}
    return poses, log
