"""tests/golden/state_keys.json: state_dict key -> shape of EVERY model variant, produced by the reference's own model/unet.py
(imported verbatim through oracle/refload.py; build container only).  tests/test_host.py checks that our constructors
yield exactly the same mapping, i.e. that any reference checkpoint loads with strict=True.

    python tools/make_state_keys.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.refload import load_reference  # noqa: E402

BASE = dict(ppm=False, aspp=False, dilation=1, no_skip=False, interpolate=False, attention=False, dec_interp=False,
            deep_supervision=False, loss_str="focal+dice", encoder="resnest50", dmg_model="siamese", type="pre", tta=False)

CASES = {
    "loc_resnest50": dict(),
    "loc_resnet50": dict(encoder="resnet50"),
    "loc_resnest50_ds_attn": dict(deep_supervision=True, attention=True),
    "loc_resnest50_dil2": dict(dilation=2),
    "loc_resnest50_dil4": dict(dilation=4),
    "loc_resnest50_noskip": dict(no_skip=True),
    "siamese": dict(type="post", dmg_model="siamese"),
    "siameseEnc": dict(type="post", dmg_model="siameseEnc"),
    "fused_ds_attn": dict(type="post", dmg_model="fused", deep_supervision=True, attention=True),
    "fusedEnc": dict(type="post", dmg_model="fusedEnc"),
    "parallel": dict(type="post", dmg_model="parallel"),
    "parallelEnc": dict(type="post", dmg_model="parallelEnc"),
    "diff": dict(type="post", dmg_model="diff"),
    # optional model parts (layers.py:6-65,138-139,175-188)
    "loc_resnet50_ppm": dict(encoder="resnet50", ppm=True),
    "loc_resnest50_aspp_dil2": dict(aspp=True, dilation=2),
    "loc_resnet50_dec_interp": dict(encoder="resnet50", dec_interp=True),
    "loc_resnet50_interpolate": dict(encoder="resnet50", interpolate=True),
    "siamese_coral": dict(type="post", dmg_model="siamese", loss_str="coral"),
    "siamese_mse_interpolate": dict(type="post", dmg_model="siamese", loss_str="mse", interpolate=True),
    "siameseEnc_ppm": dict(type="post", dmg_model="siameseEnc", ppm=True),
    "parallelEnc_aspp": dict(type="post", dmg_model="parallelEnc", aspp=True),
    "parallelEnc_ppm_interpolate": dict(type="post", dmg_model="parallelEnc", ppm=True, interpolate=True),
}


def main():
    unet, _, _ = load_reference()
    out = {}
    for name, over in CASES.items():
        ns = argparse.Namespace(**dict(BASE, **over))
        model = unet.UNetLoc(ns) if ns.type == "pre" else unet.get_dmg_unet(ns)
        out[name] = {"args": dict(BASE, **over), "keys": {k: list(v.shape) for k, v in model.state_dict().items()}}
        print(name, len(out[name]["keys"]))
    with open(os.path.join(ROOT, "tests", "golden", "state_keys.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
