"""Synthetic clouds of the reference's dataset shape (no dataset ships with the reference).

  * uniform:      torch.rand(B, N, 3), as in the reference's own demo (emd_module.py:82-83);
  * table-shaped: scene `Table`/`Cube` of robosuite_envs/envs.py:39-88 after bbox normalisation to [0,1]^3
                  (utils.py:126-143): 2048 points, classes env/cube/arm/base/gripper drawn from
                  class_distribution [0.3, 0.01, 0.4, 0.05, 0.05] (renormalised), a table-top plane, an arm
                  made of capsules between random joints, small Gaussian blobs for base/gripper/cube.
Everything is generated on the CPU from a seeded torch.Generator so that tests, bench and fixtures agree.
"""
import math

import torch

CLASS_DISTRIBUTION = [0.3, 0.01, 0.4, 0.05, 0.05]  # env, cube, arm, base, gripper (envs.py:71,87)
CLASS_COLORS = [[0.5, 0.5, 0.5], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.0]]


def uniform_clouds(b, n, seed=0, d=3):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, n, d, generator=g), torch.rand(b, n, d, generator=g)


def _capsule(g, a, b_, count, radius):
    t = torch.rand(count, 1, generator=g)
    centre = a + t * (b_ - a)
    off = torch.randn(count, 3, generator=g)
    off = off / off.norm(dim=1, keepdim=True).clamp_min(1e-6) * radius * torch.rand(count, 1, generator=g).sqrt()
    return centre + off


def table_cloud(n, g):
    """One (n, 4) cloud: xyz in [0,1]^3 and the class label as a float (dataset layout of utils.py:362-371)."""
    probs = torch.tensor(CLASS_DISTRIBUTION)
    probs = probs / probs.sum()
    labels = torch.multinomial(probs, n, replacement=True, generator=g)
    counts = torch.bincount(labels, minlength=5).tolist()
    parts, labs = [], []
    # env: table top plane
    c = counts[0]
    env = torch.stack([0.25 + 0.5 * torch.rand(c, generator=g), 0.25 + 0.5 * torch.rand(c, generator=g),
                       0.2 + 0.002 * torch.randn(c, generator=g)], dim=1)
    parts.append(env); labs.append(torch.full((c,), 0))
    # cube
    cube_c = torch.tensor([0.35, 0.35, 0.22]) + torch.rand(3, generator=g) * torch.tensor([0.3, 0.3, 0.0])
    parts.append(cube_c + 0.01 * torch.randn(counts[1], 3, generator=g)); labs.append(torch.full((counts[1],), 1))
    # arm: 3 capsules between random joints
    base_p = torch.tensor([0.5, 0.15, 0.25])
    j1 = base_p + torch.tensor([0.0, 0.05, 0.25]) + 0.05 * torch.randn(3, generator=g)
    j2 = j1 + torch.tensor([0.0, 0.2, 0.1]) + 0.08 * torch.randn(3, generator=g)
    j3 = j2 + torch.tensor([0.0, 0.15, -0.15]) + 0.08 * torch.randn(3, generator=g)
    ca = counts[2]
    split = [ca // 3, ca // 3, ca - 2 * (ca // 3)]
    arm = torch.cat([_capsule(g, base_p, j1, split[0], 0.02), _capsule(g, j1, j2, split[1], 0.02),
                     _capsule(g, j2, j3, split[2], 0.02)])
    parts.append(arm); labs.append(torch.full((ca,), 2))
    parts.append(base_p + 0.03 * torch.randn(counts[3], 3, generator=g)); labs.append(torch.full((counts[3],), 3))
    parts.append(j3 + 0.015 * torch.randn(counts[4], 3, generator=g)); labs.append(torch.full((counts[4],), 4))
    xyz = torch.cat(parts).clamp(0.0, 1.0)
    lab = torch.cat(labs).float()
    perm = torch.randperm(n, generator=g)
    return torch.cat([xyz, lab[:, None]], dim=1)[perm]


def table_clouds(b, n, seed=0, regime="independent"):
    """(pred_xyz (B,N,3), target (B,N,4)).  regime 'independent' = early training (pred is another draw),
    'noisy' = late training (pred = permuted target + N(0, 0.01))."""
    g = torch.Generator().manual_seed(seed)
    target = torch.stack([table_cloud(n, g) for _ in range(b)])
    if regime == "independent":
        pred = torch.stack([table_cloud(n, g)[:, :3] for _ in range(b)])
    elif regime == "noisy":
        pred = torch.stack([target[i, torch.randperm(n, generator=g), :3] for i in range(b)])
        pred = (pred + 0.01 * torch.randn(b, n, 3, generator=g)).clamp(0.0, 1.0)
    else:
        raise ValueError(regime)
    return pred.contiguous(), target.contiguous()


def autoencoder_batch(b, n, seed=0, regime="independent"):
    """pred (B,N,6) ~ sigmoid outputs, target (B,N,6) = xyz + rgb (Autoencoder row of SURVEY.md App. D)."""
    pred_xyz, target4 = table_clouds(b, n, seed, regime)
    g = torch.Generator().manual_seed(seed + 1000)
    colors = torch.tensor(CLASS_COLORS)
    rgb = (colors[target4[:, :, 3].long()] + 0.05 * torch.randn(b, n, 3, generator=g)).clamp(0, 1)
    target = torch.cat([target4[:, :, :3], rgb], dim=2)
    pred = torch.cat([pred_xyz, torch.rand(b, n, 3, generator=g)], dim=2)
    return pred.contiguous(), target.contiguous()


def segmenter_batch(b, n, seed=0, regime="independent", num_classes=5):
    """pred (B,N,3+C) = [sigmoid-range xyz, raw logits], target (B,N,4) (Segmenter row of SURVEY.md App. D)."""
    pred_xyz, target = table_clouds(b, n, seed, regime)
    g = torch.Generator().manual_seed(seed + 2000)
    logits = torch.randn(b, n, num_classes, generator=g)
    return torch.cat([pred_xyz, logits], dim=2).contiguous(), target


def multisegmenter_batch(b, n, seed=0):
    """pred dict {class: (B,P_c,3)} with P_c = ceil(class_distribution * n) (train.py:115-119), target (B,N,4)."""
    _, target = table_clouds(b, n, seed)
    g = torch.Generator().manual_seed(seed + 3000)
    names = ["env", "cube", "arm", "base", "gripper"]
    pred = {nm: torch.rand(b, int(math.ceil(p * n)), 3, generator=g) for nm, p in zip(names, CLASS_DISTRIBUTION)}
    labels = {nm: i for i, nm in enumerate(names)}
    return pred, target, labels
