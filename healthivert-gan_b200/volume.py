"""Batched restatement of the reference's inference driver ``eval_3d_sagittal_twostage.py`` (process_nii_files :136-241 +
run_model :46-133) on the device: per straightened volume, for every slice of the synthesis window, the iterative
three-stage synthesis (upper neighbour -> lower neighbour -> target vertebra), **stage-major**: all slices of a stage
run as one batch through the generator, slice preparation / stitching / uint8 hand-over between stages stay on the GPU
(libhv_b200.so: hv_vol_to_u8, hv_slice_id_counts, hv_slice_prepare, hv_stitch, hv_slice_finish).

The host only decides WHICH slices enter which stage (a [S, 3] int32 count table, one small D2H per volume), exactly the
control decisions the reference takes at eval:186-197, :204, :213.  ``axis=2`` walks sagittal slices ``vol[:, :, z]`` like the
reference driver; ``axis=1`` is the coronal twin (``vol[:, z, :]``, evaluation/RHLV_quantification_coronal.py:51-54) needed by
the 2.5D RHLV features.  ``synthesize`` takes and returns arrays; ``synthesize_files`` is the file-level twin of the reference
loop body (NIfTI in, NIfTI out, healthivert_gan_b200.nifti).
"""
import os

import numpy as np
import torch

from . import _lib, mask_ops, nifti
from ._lib import check, ptr


class VolumeSynthesizer:
    def __init__(self, generator, batch=64, maxheight=40):
        """generator: healthivert_gan_b200.Generator on a CUDA device, eval mode (fp32 parity or bf16 tensor-core mode)."""
        self.g = generator
        self.batch = int(batch)
        self.maxheight = int(maxheight)
        self.dev = next(generator.parameters()).device
        if self.dev.type != "cuda":
            raise _lib.HvError("VolumeSynthesizer needs the generator on a CUDA device (no CPU fallback)")

    # ---------------------------------------------------------------------------------------- helpers
    def _staging(self, numel):
        """A pinned float64 staging block of >= numel elements that no copy in flight still reads: three grow-only blocks are rotated
        (CT, label, Grad-CAM of one call), each guarded by the event recorded after its last host-to-device copy."""
        if not hasattr(self, "_staged"):
            self._staged = []
        if len(self._staged) >= 3:
            buf, ev = self._staged.pop(0)
            ev.synchronize()
        else:
            buf, ev = None, torch.cuda.Event()
        if buf is None or buf.numel() < numel:
            buf = torch.empty(int(numel), dtype=torch.float64, pin_memory=True)
        self._staged.append((buf, ev))
        return buf[:numel]

    def _to_u8_slices(self, vol, axis, scale):
        """[d0, d1, d2] float64 volume (numpy, or a torch tensor already on the device) -> uint8 slices [S, d0, ncol] along ``axis``."""
        if isinstance(vol, torch.Tensor):
            v = vol.to(self.dev, dtype=torch.float64).contiguous()
        else:
            # gather (the window is a strided view) and dtype conversion go straight into a pinned staging tensor: one host pass, then an
            # asynchronous copy at the PCIe rate (torch's pinned-memory allocator keeps the block until the copy has completed)
            vol = np.asarray(vol)
            host = self._staging(vol.size).view(vol.shape)
            np.copyto(host.numpy(), vol, casting="unsafe")
            v = host.to(self.dev, non_blocking=True)
            self._staged[-1][1].record(torch.cuda.current_stream(self.dev))
        d0, d1, d2 = v.shape
        s, ncol = (d2, d1) if axis == 2 else (d1, d2)
        out = torch.empty(s, d0, ncol, device=self.dev, dtype=torch.uint8)
        check(_lib.lib().hv_vol_to_u8(ptr(v), ptr(out), d0, d1, d2, axis, float(scale), _lib.stream()))
        return out

    @staticmethod
    def _window(vol, axis, a, b):
        """The slices a .. b (inclusive) of a volume along ``axis`` (2: vol[:, :, z], 1: vol[:, z, :])."""
        return vol[:, :, a:b + 1] if axis == 2 else vol[:, a:b + 1, :]

    def _stage(self, slices, vert_id, label_in, label_next, ct_u8, cam_u8, ratios, ct_out, label_out):
        """One run_model() per entry of ``slices`` (all with the same vertebra id), batched."""
        L = _lib.lib()
        S, h, w = label_in.shape
        dev = self.dev
        for chunk in (slices[i:i + self.batch] for i in range(0, len(slices), self.batch)):
            nb = len(chunk)
            idx = torch.tensor(chunk, device=dev, dtype=torch.int32)
            vid = torch.full((nb,), int(vert_id), device=dev, dtype=torch.int32)
            scratch = torch.empty(2 * nb * h * w, device=dev, dtype=torch.int32)
            keep = torch.empty(nb * h * w, device=dev, dtype=torch.uint8)
            meta = torch.empty(nb, 8, device=dev, dtype=torch.int32)
            f = lambda: torch.empty(nb, 1, h, w, device=dev, dtype=torch.float32)
            ct, mask, cam1m, ori = f(), f(), f(), f()
            x1, x2, hh = (torch.empty(nb, device=dev, dtype=torch.int32) for _ in range(3))
            st = _lib.stream()
            check(L.hv_slice_prepare(ptr(label_in), ptr(ct_u8), ptr(cam_u8), ptr(idx), ptr(vid), nb, h, w, self.maxheight, ptr(scratch),
                                     ptr(keep), ptr(meta), ptr(ct), ptr(mask), ptr(cam1m), ptr(ori), ptr(x1), ptr(x2), ptr(hh), st))
            ratio = torch.tensor([ratios[z] for z in chunk], device=dev, dtype=torch.float32)
            with torch.no_grad():
                _, fine_seg, _, x_stage2, _, _, pred2_h = self.g(ct, mask, cam1m, ratio)
            fake_ct, rows = mask_ops.stitch(x_stage2, ori, pred2_h, x1, x2, hh, self.maxheight, return_rows=True)
            check(L.hv_slice_finish(ptr(fake_ct), ptr(fine_seg), ptr(rows), ptr(meta), ptr(idx), ptr(vid), ptr(label_in), nb, h, w,
                                    ptr(ct_out), ptr(label_out), ptr(ct_u8), ptr(label_next), _lib.stream()))
            self.last_meta = meta

    # ---------------------------------------------------------------------------------------- driver
    @torch.no_grad()
    def synthesize(self, ct_vol, label_vol, cam_vol, vert_id, axis=2, cam_scale=255.0, return_device=False):
        """process_nii_files for ONE volume.  ct_vol / label_vol / cam_vol: [d0, d1, d2] arrays (the reference's get_fdata()
        float64 volumes; cam in [0, 1], multiplied by ``cam_scale`` like eval:181).  Returns (ct_fake, label_fake) volumes of the
        input shape (float32), zero outside the synthesis window exactly like the reference's np.zeros_like outputs."""
        vert_id = int(vert_id)
        prev_flags = (self.g.per_sample_mask, self.g.return_flow)
        self.g.per_sample_mask, self.g.return_flow = True, False     # the reference driver is batch-1: every slice has its own mask
        try:
            # the label volume is scanned whole (which slices hold the vertebra and its neighbours, eval:186-197, :204, :213) ...
            lab_all = self._to_u8_slices(label_vol, axis, 1.0)
            S, h, w = lab_all.shape
            counts = torch.empty(S, 3, device=self.dev, dtype=torch.int32)
            check(_lib.lib().hv_slice_id_counts(ptr(lab_all), S, h * w, vert_id - 1, vert_id, vert_id + 1, ptr(counts), _lib.stream()))
            cnt = counts.cpu().numpy()
            present = np.nonzero(cnt[:, 1] > 0)[0]
            full_shape = tuple(label_vol.shape)
            if not present.size:
                if return_device:
                    z = torch.zeros(full_shape, device=self.dev, dtype=torch.float32)
                    return z, z.clone()
                return np.zeros(full_shape, np.float32), np.zeros(full_shape, np.float32)
            z0, z1 = int(present.min()), int(present.max())                          # eval:186-197
            range_length = z1 - z0 + 1
            new_len = int(range_length * 4 / 5)
            new_z0 = z0 + (range_length - new_len) // 2
            new_z1 = new_z0 + new_len - 1
            center = (new_z0 + new_z1) // 2
            zs = list(range(new_z0, new_z1 + 1))
            ratios = {z - new_z0: abs(z - center) / range_length * 2 for z in zs}    # eval:202-203 (keys: index inside the window)
            upper = [z - new_z0 for z in zs if vert_id > 8 and cnt[z, 0] > 200]      # eval:204
            lower = [z - new_z0 for z in zs if vert_id < 24 and cnt[z, 2] > 200]     # eval:213
            nwin = len(zs)
            # ... but only the window's slices are ever read or written by the three stages (every run_model call works on ONE slice):
            # CT and Grad-CAM are uploaded for the window alone (a 256^3 float64 volume is 134 MB of PCIe traffic, the window ~ 15 %
            # of it), and only the window comes back
            if nwin > 0:
                ct_u8 = self._to_u8_slices(self._window(ct_vol, axis, new_z0, new_z1), axis, 1.0)
                cam_u8 = self._to_u8_slices(self._window(cam_vol, axis, new_z0, new_z1), axis, cam_scale)
                lab_a = lab_all[new_z0:new_z1 + 1].contiguous()
            ct_out = torch.zeros(nwin, h, w, device=self.dev, dtype=torch.float32)
            label_out = torch.zeros(nwin, h, w, device=self.dev, dtype=torch.float32)
            if nwin > 0:
                lab_b = lab_a.clone()
                cur, nxt = lab_a, lab_b
                for stage_slices, vid, final in ((upper, vert_id - 1, False), (lower, vert_id + 1, False), (list(range(nwin)), vert_id, True)):
                    if not stage_slices:
                        continue
                    nxt.copy_(cur)                                                   # slices outside the stage pass through
                    self._stage(stage_slices, vid, cur, nxt, ct_u8, cam_u8, ratios, ct_out if final else None,
                                label_out if final else None)
                    cur, nxt = nxt, cur
            # back to the volume's own axis order, zeros outside the window (the reference's np.zeros_like outputs)
            perm = (1, 2, 0) if axis == 2 else (1, 0, 2)
            if return_device:
                outs = []
                for win in (ct_out, label_out):
                    full = torch.zeros(full_shape, device=self.dev, dtype=torch.float32)
                    self._window(full, axis, new_z0, new_z1).copy_(win.permute(*perm))
                    outs.append(full)
                return outs[0], outs[1]
            outs = []
            if not hasattr(self, "_out_staged"):
                self._out_staged = [None, None]
            for k, win in enumerate((ct_out, label_out)):
                win = win.permute(*perm).contiguous()                                  # the volume's axis order, on the device
                if self._out_staged[k] is None or self._out_staged[k].numel() < win.numel():   # grow-only pinned blocks: D2H at the PCIe rate
                    self._out_staged[k] = torch.empty(win.numel(), dtype=torch.float32, pin_memory=True)
                host = self._out_staged[k][:win.numel()].view(win.shape)
                host.copy_(win, non_blocking=True)
                outs.append(host)
            torch.cuda.current_stream(self.dev).synchronize()
            res = []
            for host in outs:
                full = np.zeros(full_shape, np.float32)
                self._window(full, axis, new_z0, new_z1)[...] = host.numpy()           # row-wise copies, no host-side transpose
                res.append(full)
            return res[0], res[1]
        finally:
            self.g.per_sample_mask, self.g.return_flow = prev_flags

    def synthesize_files(self, ct_path, label_path, cam_path, out_ct_path, out_label_path, vert_id=None, axis=2):
        """One iteration of the reference's file loop (eval_3d_sagittal_twostage.py:153-241): read the straightened CT / label /
        Grad-CAM volumes, synthesise, write `CT_fake` / `label_fake` with the CT volume's affine.  ``vert_id`` defaults to the
        number after the last underscore of the file name (`<patient>_<vert>.nii.gz`, eval:166-167).  Outputs are float64 like
        the reference's `np.zeros_like(ct_nii.get_fdata())` volumes."""
        if vert_id is None:
            stem = os.path.basename(ct_path)
            stem = stem[:-7] if stem.endswith(".nii.gz") else os.path.splitext(stem)[0]
            vert_id = int(stem.rsplit("_", 1)[1])
        ct_img = nifti.load(ct_path)
        ct = ct_img.get_fdata()
        label = nifti.load(label_path).get_fdata()
        cam = nifti.load(cam_path).get_fdata()
        if not (ct.shape == label.shape == cam.shape and ct.ndim == 3):
            raise _lib.HvError(f"CT {ct.shape}, label {label.shape} and CAM {cam.shape} volumes must be 3-D and of one shape")
        ct_fake, label_fake = self.synthesize(ct, label, cam, vert_id, axis=axis)
        for path, vol in ((out_ct_path, ct_fake), (out_label_path, label_fake)):
            d = os.path.dirname(path)
            if d:
                os.makedirs(d, exist_ok=True)
            nifti.save(path, vol.astype(np.float64), ct_img.affine)
        return vert_id
