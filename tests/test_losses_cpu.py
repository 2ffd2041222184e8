"""Host-side logic of the fused EarthMoverDistance path (class weights, weighted CE, MSE, ratio, logging)
checked on CPU against the golden vectors of the real reference Python, with the CUDA entry points served
by the CPU oracle (test-only subclass from test_sharded_cpu)."""
import numpy as np
import pytest
import torch

from test_sharded_cpu import CpuEMD


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("tag,C", [("ae", None), ("seg", 5)])
def test_fused_host_logic_matches_reference_python(golden, tag, C):
    pred = _t(golden[f"{tag}_pred"]).requires_grad_()
    fn = CpuEMD(eps=0.005, its=50, num_classes=C)
    logged = {}
    fn.log = lambda k, v: logged.__setitem__(k, float(v.detach()))
    loss = fn(pred, _t(golden[f"{tag}_target"]))
    loss.backward()
    assert float(loss.detach()) == pytest.approx(float(golden[f"{tag}_loss"]), rel=2e-6)
    keys = ["EMD", "feature"] + (["cross_entropy", "kl_divergence"] if C else [])
    for k in keys:  # the four keys of utils.py:297-298,306-307
        assert logged[f"train_loss/{k}"] == pytest.approx(float(golden[f"{tag}_log_{k}"]), rel=1e-5)
    np.testing.assert_allclose(pred.grad.numpy(), golden[f"{tag}_grad"], rtol=2e-5, atol=1e-10)


def test_log_attribute_protocol_is_optional():
    """train.py:161 assigns `.log` after construction; without it the loss must still work."""
    from pointcloud_b200 import synth
    pred, target = synth.autoencoder_batch(1, 128, seed=1)
    fn = CpuEMD(0.005, 10)
    assert torch.isfinite(fn(pred, target))
    assert fn.feature_weight == 0.1 and fn.eps == 0.005 and fn.iterations == 10 and fn.C is None  # ctor surface, utils.py:246-251
