"""Import the REAL /root/reference/pointcloud_vision/train.py in this container (test infrastructure).

train.py:1-16 imports Lightning, the tensorboard logger, the model zoo (-> pointnet2_ops, pytorch3d), pc_encoder
(-> robosuite_envs, gymnasium) and robosuite_envs.envs (-> robosuite, MuJoCo); none of those is installed here and none
of them is on the loss path.  They are replaced by the thinnest possible stand-ins so that the reference's own
`Lit` (train.py:19-68) and `create_model` (train.py:71-163) run UNMODIFIED:

  * lightning.pytorch.LightningModule -> torch.nn.Module + a `log(name, value)` that records what it is given;
  * robosuite / gymnasium / the sibling modules of robosuite_envs -> permissive empty modules, so that
    robosuite_envs/envs.py executes and defines the real `cfg_scene` table (envs.py:31-137);
  * pointnet2_ops.furthest_point_sample -> the CPU FPS oracle; pytorch3d / emd -> as in make_golden.install_stubs.
Nothing here is imported by the product package or on the GPU box.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


class _Anything:
    """Stands for any class / function / constant of an absent third-party module."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __mro_entries__(self, bases):  # `class X(Anything())` -> plain object subclass
        return (object,)


class _PermissiveModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


def _permissive(name, **attrs):
    m = _PermissiveModule(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class LightningModuleStub(torch.nn.Module):
    """What `Lit` uses of pl.LightningModule: nn.Module behaviour + self.log (train.py:34,40)."""

    def __init__(self):
        super().__init__()
        self.logged = {}

    def log(self, name, value, *a, **k):
        self.logged[name] = value


def available():
    return os.path.exists(os.path.join(REF, "pointcloud_vision", "train.py"))


def install():
    """Returns the imported reference modules (train, utils).  Idempotent."""
    if "pointcloud_vision.train" in sys.modules:
        return sys.modules["pointcloud_vision.train"], sys.modules["pointcloud_vision.utils"]
    sys.path.insert(0, HERE)
    import make_golden
    make_golden.install_stubs()  # emd, pytorch3d, the pointcloud_vision package shell, CPU redirection of .cuda()
    import oracle
    sys.modules["pytorch3d.ops"].sample_farthest_points = lambda pts, K=50, **k: (
        lambda idx: (pts.gather(1, idx.unsqueeze(-1).expand(-1, -1, pts.shape[2])), idx))(
        torch.from_numpy(oracle.fps(pts[:, :, :3], K)).long())
    p2o = types.ModuleType("pointnet2_ops")
    p2u = types.ModuleType("pointnet2_ops.pointnet2_utils")
    p2u.furthest_point_sample = lambda xyz, npoint: torch.from_numpy(oracle.fps(xyz, npoint))
    p2o.pointnet2_utils = p2u
    sys.modules.update({"pointnet2_ops": p2o, "pointnet2_ops.pointnet2_utils": p2u})
    _permissive("pointnet2_ops.pointnet2_modules")

    pl = _permissive("lightning.pytorch", LightningModule=LightningModuleStub)
    _permissive("lightning", pytorch=pl)
    _permissive("pytorch_lightning")
    _permissive("pytorch_lightning.loggers")
    _permissive("robosuite")
    _permissive("robosuite.controllers", load_controller_config=lambda **k: {})
    for sub in ("wrappers", "utils", "utils.camera_utils", "utils.transform_utils", "utils.mjcf_utils", "models", "models.objects"):
        _permissive("robosuite." + sub)
    for name in ("gymnasium", "gymnasium.spaces", "gymnasium.envs", "gymnasium.envs.registration", "open3d", "matplotlib", "matplotlib.pyplot",
                 "stable_baselines3", "sb3_contrib"):
        _permissive(name)
    # robosuite_envs: run the REAL envs.py (it holds cfg_scene), with its sibling modules stubbed
    pkg = types.ModuleType("robosuite_envs")
    pkg.__path__ = [os.path.join(REF, "robosuite_envs")]
    sys.modules["robosuite_envs"] = pkg
    for sub in ("base_env", "encoders", "sensors", "utils"):
        _permissive("robosuite_envs." + sub)
    import pointcloud_vision.utils as ref_utils
    import pointcloud_vision.pc_encoder  # noqa: F401  first, as the reference package does: train.py <-> pc_encoder.py import each other
    import pointcloud_vision.train as ref_train
    return ref_train, ref_utils
