"""Generate tests/golden/train_golden.npz by running the REAL reference caller: `create_model` + `Lit.training_step`
(/root/reference/pointcloud_vision/train.py:19-35,71-163) with the reference's own models and loss classes (third-party
imports stubbed as described in ref_train_stubs.py / make_golden.py).  Recorded per model type: the batch, the model's
prediction, the loss, everything `Lit.log` / `loss_fn.log` received, and d loss / d prediction.
The GPU box has no /root/reference: tests/test_dropin_gpu.py replays these vectors through the CUDA losses.
Usage: python tests/golden/make_train_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_train_stubs  # noqa: E402


def main():
    train, _ = ref_train_stubs.install()
    from test_dropin_train_cpu import _batch
    out = {}
    for model_type, backbone in (("Autoencoder", "PointNet2"), ("Segmenter", "PointNet"), ("MultiSegmenter", "PointNet")):
        torch.manual_seed(0)
        lit, _ = train.create_model(model_type, backbone, "Cube")
        lit.train()
        x, y = _batch(model_type, 2, 2048)
        kept = {}
        inner = lit.model.forward

        def forward(inp, _inner=inner, _kept=kept):
            pred = _inner(inp)
            for v in (pred.values() if isinstance(pred, dict) else [pred]):
                v.retain_grad()
            _kept["pred"] = pred
            return pred

        lit.model.forward = forward
        torch.manual_seed(1)
        loss = lit.training_step((x, y), 0)
        loss.backward()
        tag = model_type
        out[f"{tag}_y"] = y.numpy()
        out[f"{tag}_loss"] = np.float32(loss.item())
        pred = kept["pred"]
        for name, v in (pred.items() if isinstance(pred, dict) else [("", pred)]):
            out[f"{tag}_pred_{name}"] = v.detach().numpy()
            out[f"{tag}_grad_{name}"] = v.grad.numpy()
        for k, v in lit.logged.items():
            out[f"{tag}_log_{k.replace('/', '.')}"] = np.float32(float(v))
        print(tag, float(loss), {k: float(v) for k, v in lit.logged.items()})
    path = os.path.join(HERE, "train_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
