"""Generates tests/golden/*.npz by running the UNMODIFIED reference kernels (oracle/_ref, built from
/root/reference/src by oracle/Makefile) on a B200:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

then the .npz files are copied from gpurun_out/golden/ into tests/golden/ and committed.  Every file
holds the inputs and the reference's outputs, so the CPU tests need neither a GPU nor /root/reference.

It also pins the CPU oracle at the largest shapes the reference can launch (w, h <= 1024) and
writes the comparison to <out>/ORACLE_PIN_REPORT.json.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle as ko  # noqa: E402
from oracle import ref_gpu as ref  # noqa: E402
from kangaroo_b200.synth import stereo_pair  # noqa: E402


def main(out_dir: str) -> None:
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.Generator(np.random.PCG64(20261017))
    save = lambda name, **kw: np.savez_compressed(os.path.join(out_dir, name + ".npz"), **kw)  # noqa: E731

    # ---- census: all windows x both input types (random + low-range image to force ties)
    img_u8 = rng.integers(0, 256, (40, 48), dtype=np.uint8)
    img_tie = rng.integers(0, 4, (40, 48), dtype=np.uint8)
    img_f32 = rng.random((40, 48), dtype=np.float32)
    cen = {}
    for nm, im in (("u8", img_u8), ("tie", img_tie), ("f32", img_f32)):
        for win in (0, 1, 2):
            cen[f"out_{nm}_{win}"] = ref.census(im, win)
    save("census", img_u8=img_u8, img_tie=img_tie, img_f32=img_f32, **cen)

    # ---- a small textured pair used by the matching stages
    L, R, gt = stereo_pair(64, 24, 16, config=90)
    cl = {win: ref.census(L, win) for win in (0, 1, 2)}
    cr = {win: ref.census(R, win) for win in (0, 1, 2)}

    cs = {"left": L, "right": R}
    for md in (16, -16, 5):
        cs[f"disp_{md}"] = ref.census_stereo(cl[0], cr[0], md)
    save("census_stereo", **cs)

    csv = {"left": L, "right": R}
    for win in (0, 1, 2):
        for sd in (-1.0, 1.0):
            csv[f"f32_w{win}_sd{int(sd)}"] = ref.census_stereo_volume(cl[win], cr[win], 16, sd, np.float32, depth=18,
                                                                     fill=7.0)
        csv[f"u16_w{win}"] = ref.census_stereo_volume(cl[win], cr[win], 16, -1.0, np.uint16, depth=16, fill=9)
    save("census_stereo_volume", **csv)

    # ---- SGM: every flag combination on a float volume; CostVolElem/uchar instantiation once
    Ls, Rs, _ = stereo_pair(40, 24, 12, config=91)
    volc = ref.census_stereo_volume(ref.census(Ls, 0), ref.census(Rs, 0), 12, -1.0)
    left_f = Ls.astype(np.float32) * np.float32(1.0 / 255.0)
    sg = {"left_u8": Ls, "left_f32": left_f, "volc": volc}
    for hz in (0, 1):
        for vt in (0, 1):
            for rv in (0, 1):
                sg[f"H_h{hz}v{vt}r{rv}"] = ref.sgm(volc, left_f, 12, 0.01, 0.02, hz, vt, rv)
    sg["H_md7"] = ref.sgm(volc, left_f, 7, 0.05, 0.3, 1, 1, 1)
    volr = rng.random((12, 24, 40), dtype=np.float32)
    sg["volc_rand"] = volr
    sg["H_rand"] = ref.sgm(volr, left_f, 12, 0.1, 0.4, 1, 1, 1)
    elem = np.zeros((12, 24, 40), ko.COSTVOLELEM)
    elem["n"] = rng.integers(0, 4, elem.shape)
    elem["sum"] = rng.random(elem.shape, dtype=np.float32) * 3
    sg["volc_elem_n"] = elem["n"]
    sg["volc_elem_sum"] = elem["sum"]
    sg["H_elem"] = ref.sgm(elem, Ls, 12, 1.0, 8.0, 1, 1, 1)
    save("sgm", **sg)

    # ---- WTA (reference kernel is unguarded: shapes are multiples of 32)
    wt = {}
    vf = rng.random((20, 32, 64), dtype=np.float32)
    vf[:, ::3, ::5] = 0.25  # ties: first minimum must win
    wt["vol_f32"] = vf
    wt["disp_f32_f32"] = ref.costvol_minimum(vf, 20, np.float32)
    wt["disp_i8_f32"] = ref.costvol_minimum(vf, 13, np.int8)
    for nm, dt in (("i32", np.int32), ("u32", np.uint32), ("u16", np.uint16), ("u8", np.uint8)):
        v = rng.integers(0, 50, (20, 32, 64)).astype(dt)
        if dt == np.int32:
            v -= 25
        wt[f"vol_{nm}"] = v
        wt[f"disp_i8_{nm}"] = ref.costvol_minimum(v, 20, np.int8)
    wt["disp_f32_u16"] = ref.costvol_minimum(wt["vol_u16"], 20, np.float32)
    el = np.zeros((9, 32, 64), ko.COSTVOLELEM)
    el["n"] = rng.integers(0, 3, el.shape)
    el["sum"] = rng.random(el.shape, dtype=np.float32)
    wt["elem_n"], wt["elem_sum"] = el["n"], el["sum"]
    wt["disp_elem"] = ref.costvol_minimum_elem(el)
    save("costvol_minimum", **wt)

    # ---- WTA + parabola (guarded kernel: any shape); depth = maxDisp + 1 keeps bestd+1 in bounds
    sp = {}
    vs = rng.random((17, 20, 50), dtype=np.float32)
    sp["vol"] = vs
    sp["disp_sd-1"] = ref.costvol_minimum_subpix(vs, 16, -1.0)
    sp["disp_sd1"] = ref.costvol_minimum_subpix(vs, 16, 1.0)
    Hs = sg["H_h1v1r1"]
    Hpad = np.concatenate([Hs, np.zeros((1,) + Hs.shape[1:], np.float32)], 0)
    sp["vol_sgm"] = Hpad
    sp["disp_sgm"] = ref.costvol_minimum_subpix(Hpad, 12, -1.0)
    save("costvol_minimum_subpix", **sp)

    # ---- SAND 5x5 subpixel refinement
    Lr, Rr, gtr = stereo_pair(72, 40, 16, config=92)
    dr8 = np.clip(gtr + rng.integers(-1, 2, gtr.shape), 0, 255).astype(np.uint8)
    save("dense_stereo_subpixel_refine", left=Lr, right=Rr, disp=dr8,
         out=ref.dense_stereo_subpixel_refine(dr8, Lr, Rr))

    # ---- left-right check
    dl = (rng.random((20, 50), dtype=np.float32) * 12).astype(np.float32)
    drr = (rng.random((20, 50), dtype=np.float32) * 12).astype(np.float32)
    dl[3, 4] = np.nan
    drr[::4, ::3] = np.nan
    drr[1::4, 1::3] = np.inf
    lr = {"dl": dl, "dr": drr}
    lr["f32_sd-1_0.5"] = ref.left_right_check_f32(dl, drr, -1.0, 0.5)
    lr["f32_sd1_4"] = ref.left_right_check_f32(dl, drr, 1.0, 4.0)
    dli = rng.integers(0, 6, (20, 200)).astype(np.int8)
    dri = rng.integers(0, 3, (20, 200)).astype(np.int8)
    lr["dli"], lr["dri"] = dli, dri
    lr["i8_sd-1_0"] = ref.left_right_check_i8(dli, dri, -1, 0)
    lr["i8_sd1_2"] = ref.left_right_check_i8(dli, dri, 1, 2)
    save("left_right_check", **lr)

    # ---- whole path on one pair, stage by stage, as applications/stereo2/main.cpp:375-454 runs it
    Lp, Rp, gtp = stereo_pair(96, 64, 32, config=93)
    imgf = [a.astype(np.float32) * np.float32(1.0 / 255.0) for a in (Lp, Rp)]
    pipe = {"left": Lp, "right": Rp, "gt": gtp}
    for win in (0, 2):
        c0, c1 = ref.census(imgf[0], win), ref.census(imgf[1], win)
        v0 = ref.census_stereo_volume(c0, c1, 32, -1.0, depth=33)
        v1 = ref.census_stereo_volume(c1, c0, 32, 1.0, depth=33)
        H = ref.sgm(v0, imgf[0], 32, 0.01, 0.02, 1, 1, 1)
        H[32] = 0  # slice 32 exists only to keep the parabola's bestd+1 read in bounds
        d0 = ref.costvol_minimum_subpix(H, 32, -1.0)
        d1 = ref.costvol_minimum_subpix(v1, 32, 1.0)
        d1c = ref.left_right_check_f32(d1, d0, 1.0, 1.0)
        d0c = ref.left_right_check_f32(d0, d1c, -1.0, 1.0)
        pipe.update({f"w{win}_census0": c0, f"w{win}_H": H[:32], f"w{win}_disp0": d0, f"w{win}_disp1": d1,
                     f"w{win}_disp0_lr": d0c})
    save("pipeline", **pipe)

    # ---- pin the CPU oracle at the largest reference-runnable shapes
    report = {"device": None, "cases": []}
    import torch
    report["device"] = torch.cuda.get_device_name(0)
    for (w, h, D, cfg) in ((640, 480, 64, 1), (1024, 720, 128, 2), (1024, 375, 128, 3), (1024, 1024, 256, 4)):
        Lb, Rb, _ = stereo_pair(w, h, D, config=cfg)
        lf = Lb.astype(np.float32) * np.float32(1.0 / 255.0)
        case = {"w": w, "h": h, "D": D}
        t0 = time.time()
        for win in (0, 2):
            rc0, rc1 = ref.census(Lb, win), ref.census(Rb, win)
            oc0 = ko.census(Lb, win)
            case[f"census_w{win}_bitexact"] = bool((rc0 == oc0).all())
            rv = ref.census_stereo_volume(rc0, rc1, D, -1.0)
            ov = ko.census_stereo_volume(rc0, rc1, D, -1.0)
            case[f"volume_w{win}_bitexact"] = bool((rv == ov).all())
            if win == 0:
                rH = ref.sgm(rv, lf, D, 0.01, 0.02, 1, 1, 1)
                oH = ko.sgm(rv, lf, D, 0.01, 0.02, 1, 1, 1)
                den = np.maximum(np.abs(rH), 1e-30)
                case["sgm_max_rel"] = float((np.abs(rH - oH) / den).max())
                xs = np.arange(w)[None, None, :]
                ds = np.arange(D)[:, None, None]
                case["sgm_zero_region_exact"] = bool((rH[np.broadcast_to(ds > xs, rH.shape)] == 0).all()
                                                     and (oH[np.broadcast_to(ds > xs, oH.shape)] == 0).all())
                rHp = np.concatenate([rH, np.zeros((1, h, w), np.float32)], 0)
                oHp = np.concatenate([oH, np.zeros((1, h, w), np.float32)], 0)
                rd = ref.costvol_minimum_subpix(rHp, D, -1.0)
                od, om = ko.costvol_minimum_subpix(oHp, D, -1.0)
                case["wta_int_agree"] = float((np.rint(rd) == np.rint(od)).mean())
                both = np.isfinite(rd) & np.isfinite(od) & (om == 0) & (np.rint(rd) == np.rint(od))
                case["subpix_max_abs"] = float(np.abs(rd - od)[both].max())
        case["seconds"] = time.time() - t0
        report["cases"].append(case)
        print(case, flush=True)
    with open(os.path.join(out_dir, "ORACLE_PIN_REPORT.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
