"""SURVEY 8f N4 on the GPU: a reference-format checkpoint through the product.

The reference's own model code (oracle/_ref/py, staged verbatim at build time) on the reference's own CUDA
extension (oracle/_ref/nofma) builds a model, gives it trained-looking BatchNorm state and saves
``model.state_dict()`` to model.pth (lib/solver.py:142, 325).  The mirror loads that file with
``strict=True`` (scripts/eval.py:183-188) and evaluates the same chunks:

  * torch.nn formulation on this library's op kernels (fused.ENABLED = False): logits IDENTICAL to the
    reference's (bit for bit: same indices, same un-contracted interpolation, same torch layers);
  * fused tcgen05 path (the product's default): logits within the TF32 class bound of the fp32 logits
    (4x the deviation of the reference's own formulation under its default TF32, floor 2e-2 of the max
    magnitude) and the same arg-max class for >= 98 % of the points (random-initialised weights: tiny margins).

Runs in its own process (it installs the reference's ``pointnet2._ext`` seam).  Exit code 0 = passed.
    python tools/checkpoint_roundtrip.py [--msg] [--multiview]
"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    msg, mv = "--msg" in sys.argv, "--multiview" in sys.argv
    c0 = 131 if mv else 6
    import bench
    from oracle import build_ref
    from pointnet2.scannet_b200 import fused, pointnet2_semseg as mine, synthetic
    ref_ext = build_ref.load_ref("nofma")
    semseg, _ = bench.load_reference_python(ref_ext)
    dev = "cuda:0"
    torch.manual_seed(3)
    ref_model = semseg.get_model(num_classes=20, is_msg=msg, input_channels=c0, use_xyz=True, bn=True).to(dev)
    g = torch.Generator(device=dev).manual_seed(4)
    for m in ref_model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.normal_(0, 0.2, generator=g), m.running_var.uniform_(0.5, 2.0, generator=g)
            m.weight.data.uniform_(0.5, 1.5, generator=g), m.bias.data.normal_(0, 0.1, generator=g)
            m.num_batches_tracked.fill_(1234)
    path = os.path.join(tempfile.mkdtemp(), "model.pth")
    torch.save(ref_model.state_dict(), path)
    mirror = mine.get_model(20, is_msg=msg, input_channels=c0).to(dev)
    res = mirror.load_state_dict(torch.load(path, map_location=dev), strict=True)
    assert not res.missing_keys and not res.unexpected_keys, res
    x, _, _ = synthetic.make_batch(2, 4, 8192, use_multiview=mv, use_color=not mv)
    pc = torch.from_numpy(x).to(dev)
    ref_model.eval(), mirror.eval()

    def logits(model, tf32, fused_on):
        saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, fused.ENABLED)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
        fused.ENABLED = fused_on
        try:
            with torch.no_grad():
                return model(pc)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, fused.ENABLED = saved

    truth = logits(ref_model, False, False)
    ref_tf32 = logits(ref_model, True, False)
    unfused = logits(mirror, False, False)
    prod = logits(mirror, True, True)
    torch.cuda.synchronize()
    assert truth.shape == prod.shape == (4, 8192, 20)
    assert torch.equal(unfused, truth), "mirror (torch.nn formulation on this library's ops) != reference"
    scale = float(truth.abs().max())
    e_ref, e_prod = float((ref_tf32 - truth).abs().max()) / scale, float((prod - truth).abs().max()) / scale
    agree = float((prod.argmax(-1) == truth.argmax(-1)).float().mean())
    agree_ref = float((ref_tf32.argmax(-1) == truth.argmax(-1)).float().mean())
    print("checkpoint round trip %s C0=%d: unfused == reference bit for bit; fused deviation %.3e of max (reference under "
          "TF32: %.3e); arg-max agreement %.4f (reference under TF32: %.4f)" % (
              "MSG" if msg else "SSG", c0, e_prod, e_ref, agree, agree_ref))
    assert e_prod <= max(4 * e_ref, 2e-2), (e_prod, e_ref)
    # (random-initialised weights: the class margins are tiny, so the arg-max is a sensitive probe)
    assert agree >= min(0.98, agree_ref - 0.02), (agree, agree_ref)


if __name__ == "__main__":
    main()
