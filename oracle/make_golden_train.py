"""Generate tests/golden/train_step_n2.npz (and, with `--n 16 --steps 1`, train_step_n16.npz: BASELINE.json config 4's batch
size) by running the UNMODIFIED reference Pix2PixModel on CPU (build container only).

TEST INFRASTRUCTURE.  Run from the repo root:  python -m oracle.make_golden_train [--n N] [--steps K]
Two optimize_parameters() steps (models/pix2pix_model.py:356-382) on synthetic_train_batch(n=2, seed=7) with the shared
synthetic generator / discriminator state_dicts.  Stored per step: the 12 loss_names; after step 1 additionally, for every
parameter of G, D_1, D_2, D_3: the gradient L2 norm, 8 gradient probes and 8 probes of the updated parameter; the
spectral-norm u buffers' probes; BatchNorm running statistics probes.
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.simplefilter("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refshim, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
PROBES = np.random.Generator(np.random.PCG64(5)).random(8)


def probe(t):
    t = t.detach().double().reshape(-1)
    idx = torch.from_numpy((PROBES * t.numel()).astype(np.int64))
    return t[idx].numpy()


def build_reference(n):
    refshim.install()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from models.pix2pix_model import Pix2PixModel
        opt = synth.train_options()
        torch.manual_seed(0)
        m = Pix2PixModel(opt)
        m.setup(opt)
    m.netG.load_state_dict(synth.synthetic_generator_state_dict())
    for k, net in enumerate((m.netD_1, m.netD_2, m.netD_3), start=1):
        net.load_state_dict(synth.synthetic_discriminator_state_dict(seed=k))
    # the reference's Generator(params, True) calls .cuda() on the mask / ratio planes: no-op shim on this CPU host
    return m


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    n = args.n
    torch.set_num_threads(os.cpu_count() or 1)
    m = build_reference(n)
    batch = synth.synthetic_train_batch(n=n, seed=7)
    out = {}
    for step in range(1, args.steps + 1):
        m.set_input(batch)
        m.optimize_parameters()
        out[f"losses_step{step}"] = np.array([float(getattr(m, "loss_" + k)) for k in m.loss_names])
        if step == 1:
            out["loss_names"] = np.array(m.loss_names)
            for tag, net in (("G", m.netG), ("D_1", m.netD_1), ("D_2", m.netD_2), ("D_3", m.netD_3)):
                names, gnorm, gprobe, pprobe = [], [], [], []
                for name, p in net.named_parameters():
                    names.append(name)
                    gnorm.append(float(p.grad.double().norm()))
                    gprobe.append(probe(p.grad))
                    pprobe.append(probe(p))
                out[f"{tag}_names"] = np.array(names)
                out[f"{tag}_grad_norm"] = np.array(gnorm)
                out[f"{tag}_grad_probe"] = np.stack(gprobe)
                out[f"{tag}_param_probe"] = np.stack(pprobe)
            sd = m.netG.state_dict()
            out["u_probe"] = np.stack([probe(v) for k, v in sd.items() if k.endswith("weight_u")])
            out["bn_running"] = np.stack([probe(v.float()) for k, v in m.netD_1.state_dict().items() if "running" in k])
            out["fake_B_probe"] = probe(m.fake_B)
            out["pred_h"] = np.concatenate([m.pred1_h.detach().numpy().reshape(-1), m.pred2_h.detach().numpy().reshape(-1)])
        print("step", step, dict(zip(m.loss_names, np.round(out[f"losses_step{step}"], 5))))
    np.savez_compressed(os.path.join(GOLD, f"train_step_n{n}.npz"), **out)
    print(f"wrote train_step_n{n}.npz")


if __name__ == "__main__":
    main()
