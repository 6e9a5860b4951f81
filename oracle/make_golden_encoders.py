"""Generate tests/golden/encoders.npz by running the UNMODIFIED reference encoders (authoring container only).

Usage:  python oracle/make_golden_encoders.py        (needs /root/reference; CPU, seconds)

Test infrastructure, not product code.  The reference's `Filter.Filter` (`Filter.py:132-228`) and
`networks.define_G` (`networks.py:35-60`) are built for the small cases of
`pifu_b200.synthetic.ENCODER_CASES`, filled with key-addressed seeded parameters
(`synthetic.fill_state`), put in eval mode and run on a seeded input; outputs and the state_dict
signature are stored.  tests/test_encoders_cpu.py rebuilds the same parameters on this package's
modules from the seeds - no weights are committed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("PIFU_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "encoders.npz")


def main():
    from pifu_b200 import synthetic as syn
    sys.path.insert(0, REF)
    import Filter as ref_filter            # the reference's module
    import networks as ref_networks
    torch.set_grad_enabled(False)
    torch.set_num_threads(1)               # one summation order
    out = {}
    for seed, (name, (kind, args, shape)) in enumerate(sorted(syn.ENCODER_CASES.items()), start=11):
        net = ref_filter.Filter(*args) if kind == "filter" else ref_networks.define_G(*args)
        syn.fill_state(net, seed)
        net.eval()
        x = syn.encoder_input(shape, seed + 100)
        y = net(x)
        out[name + "_sig"] = np.frombuffer(bytes.fromhex(syn.state_signature(net)), dtype=np.uint8)
        out[name + "_seed"] = np.int64(seed)
        if kind == "filter":
            feats, normx = y
            for i, f in enumerate(feats):
                out["%s_out%d" % (name, i)] = f.numpy()
            out[name + "_normx"] = normx.numpy()
            print(name, [tuple(f.shape) for f in feats], tuple(normx.shape), float(feats[-1].abs().mean()))
        else:
            out[name + "_out0"] = y.numpy()
            print(name, tuple(y.shape), float(y.abs().mean()))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
