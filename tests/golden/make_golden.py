"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on small
seeded inputs.  Run in the CPU container where the reference is mounted:

    python tests/golden/make_golden.py

The reference modules are loaded by file path (no install, no bytecode written).  The per-pixel
loops are executed exactly as the scripts write them (src/3-capture_decode.py:95-100,
src/4-triangulate.py:63-71).
"""
import importlib.util
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


dec = load("_ref_decode_codes", "scanner/grayCode/decode_codes.py")
gen = load("_ref_generate_codes", "scanner/grayCode/generate_codes.py")
tri = load("_ref_triangulate", "scanner/triangulation/triangulate.py")


sys.path.insert(0, HERE)
from make_golden_inputs import adversarial_stack  # noqa: E402


def decode_case(seed, n, h, w, runs):
    rng = np.random.default_rng(seed)
    stacks = [adversarial_stack(rng, n, h, w) for _ in range(runs)]
    cache_h, cache_v, lds, lgs = [], [], [], []
    for st in stacks:
        imgs = st.astype(np.float64)                          # src/3-capture_decode.py:69
        with np.errstate(all="ignore"):
            l_d, l_g = dec.get_direct_indirect(imgs)
            hc, vc = dec.get_codes(imgs)
        cache_h.append(hc); cache_v.append(vc); lds.append(l_d); lgs.append(l_g)
    best_h = np.max(cache_h, axis=0)                          # :95-96
    best_v = np.max(cache_v, axis=0)
    h_pixels = np.array([dec.gray_to_decimal(best_h[:, y, x]) for y in range(h) for x in range(w)]).reshape(h, w)
    v_pixels = np.array([dec.gray_to_decimal(np.flip(best_v[:, y, x])) for y in range(h) for x in range(w)]).reshape(h, w)
    return dict(stacks=np.stack(stacks), h_codes=np.stack(cache_h), v_codes=np.stack(cache_v),
                L_d=np.stack(lds), L_g=np.stack(lgs), h_pixels=h_pixels, v_pixels=v_pixels)


def main():
    out = {}
    for name, (seed, n, h, w, runs) in {"n3": (11, 3, 12, 16, 1), "n5": (12, 5, 12, 20, 2),
                                        "n10": (13, 10, 10, 16, 2), "n11": (14, 11, 8, 16, 1),
                                        "n12": (15, 12, 8, 16, 1)}.items():
        for k, v in decode_case(seed, n, h, w, runs).items():
            out[f"decode_{name}_{k}"] = v
    # eps variants of get_is_lit (decode_codes.py:125)
    rng = np.random.default_rng(21)
    st = adversarial_stack(rng, 4, 10, 16).astype(np.float64)
    with np.errstate(all="ignore"):
        l_d, l_g = dec.get_direct_indirect(st)
        out["eps_stack"] = st.astype(np.uint8)
        for eps in (0, 2, 0.5, 3.25):
            hc, vc = dec.get_is_lit(st, l_d, l_g, eps=eps)
            out[f"eps_{eps}_h"] = hc; out[f"eps_{eps}_v"] = vc
    # pattern generator (generate_codes.py)
    for (pw, ph) in ((16, 8), (32, 32), (24, 10), (8, 16)):
        out[f"patterns_{pw}x{ph}"] = gen.get_image_sequence(gen.get_gray_codes(pw, ph), pw, ph)
        out[f"graycodes_{pw}x{ph}"] = gen.get_gray_codes(pw, ph)
    # ideal sequence decodes to the identity map
    np.savez_compressed(os.path.join(HERE, "decode_golden.npz"), **out)

    # ---- triangulation with the shipped intrinsics + a synthetic stereo pose
    import cv2
    cal = os.path.join(REF, "data", "calib_results")
    cam_mtx = np.load(os.path.join(cal, "cam_1080", "cam_mtx.npy"))
    cam_dist = np.load(os.path.join(cal, "cam_1080", "cam_dist.npy"))
    cam2_mtx = np.load(os.path.join(cal, "cam_1440", "cam_mtx.npy"))
    cam2_dist = np.load(os.path.join(cal, "cam_1440", "cam_dist.npy"))
    proj_mtx = np.load(os.path.join(cal, "proj", "proj_mtx.npy"))
    proj_dist = np.load(os.path.join(cal, "proj", "proj_dist.npy"))
    R, _ = cv2.Rodrigues(np.array([0.02, -0.15, 0.01]))
    T = np.array([[0.20], [0.02], [0.03]])
    rng = np.random.default_rng(31)
    tout = dict(R=R, T=T)
    cases = {"a": ((64, 36), (1920, 1080), (1024, 768), cam_mtx, cam_dist),
             "b": ((48, 40), (1024, 768), (1024, 768), cam_mtx, cam_dist),
             "c": ((40, 30), (1920, 1080), (1920, 1080), cam2_mtx, cam2_dist)}   # c: projector out of its calibrated domain
    for name, (cam_size, proj_size, calib_size, cm, cd) in cases.items():
        w, h = cam_size
        n_max = 2048 if proj_size[0] > 1024 else 1024
        h_pixels = rng.integers(-1, n_max, (h, w))
        v_pixels = rng.integers(-1, n_max, (h, w))
        h_pixels[rng.random((h, w)) < .2] = -1
        white = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
        pm = proj_mtx.copy()
        t = tri.Triangulate(h_pixels, v_pixels, cam_size, cm.copy(), cd, proj_size, calib_size, pm, proj_dist, R, T)
        cam_pts, proj_pts, colors = t.get_cam_proj_pts(white)
        # spread the camera points over the full sensor so the distortion terms matter
        cam_full = cam_pts * np.array([[1920 / w, 1080 / h]], dtype=np.float32)
        with np.errstate(all="ignore"):
            pts = t.triangulate(cam_full, proj_pts)
            pts_small = t.triangulate(cam_pts, proj_pts)
        fp, fc = t.filter_3d_pts(pts, colors, .5)
        fp2, fc2 = t.filter_3d_pts(pts, colors, 1.25)
        und_cam = cv2.undistortPoints(np.expand_dims(cam_full, 1), cm, cd, R=R)[:, 0]
        und_proj = cv2.undistortPoints(np.expand_dims(proj_pts, 1), pm, proj_dist)[:, 0]
        tout.update({f"{name}_h_pixels": h_pixels, f"{name}_v_pixels": v_pixels, f"{name}_white": white,
                     f"{name}_cam_size": np.array(cam_size), f"{name}_proj_size": np.array(proj_size),
                     f"{name}_calib_size": np.array(calib_size), f"{name}_cam_mtx": cm, f"{name}_cam_dist": cd,
                     f"{name}_proj_mtx_in": proj_mtx, f"{name}_proj_mtx_scaled": pm, f"{name}_proj_dist": proj_dist,
                     f"{name}_cam_pts": cam_pts, f"{name}_proj_pts": proj_pts, f"{name}_colors": colors,
                     f"{name}_cam_full": cam_full, f"{name}_pts": pts, f"{name}_pts_small": pts_small,
                     f"{name}_filt_pts": fp, f"{name}_filt_colors": fc, f"{name}_filt125_pts": fp2,
                     f"{name}_filt125_colors": fc2, f"{name}_und_cam": und_cam, f"{name}_und_proj": und_proj})
    np.savez_compressed(os.path.join(HERE, "triangulate_golden.npz"), **tout)
    # ---- capture front end: to_gray (decode_codes.py:70-87) and remove_bad_images (:34-68)
    rng = np.random.default_rng(41)
    fout = {}
    bgr = rng.integers(0, 256, (3, 20, 36, 3)).astype(np.uint8)
    bgr[0, 0, :6] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 2, 3]]
    fout["gray_bgr"] = bgr
    fout["gray_out"] = dec.to_gray(bgr.astype(np.float64))                  # float64 (N,H,W)
    fout["gray_float_in"] = bgr.astype(np.float64) + 0.75                   # astype(uint8) truncates
    fout["gray_float_out"] = dec.to_gray(fout["gray_float_in"])
    # a "video": stable pattern frames separated by blended transition frames, with sensor noise
    def video(seed, n_patterns, holds):
        r = np.random.default_rng(seed)
        pats = [r.integers(0, 256, (12, 18, 3)).astype(np.float64) for _ in range(n_patterns)]
        frames = []
        for k, pat in enumerate(pats):
            for _ in range(holds[k % len(holds)]):
                frames.append(np.clip(pat + r.integers(-3, 4, pat.shape), 0, 255))
            if k + 1 < n_patterns:
                alpha = r.uniform(0.3, 0.7)
                frames.append(np.clip(alpha * pat + (1 - alpha) * pats[k + 1], 0, 255))
        return np.floor(np.array(frames))
    for name, (seed, n_p, holds) in {"v1": (51, 6, [2]), "v2": (52, 7, [2, 3]), "v3": (53, 5, [1, 2])}.items():
        frames = video(seed, n_p, holds)
        fout[f"video_{name}"] = frames.astype(np.uint8)
        fout[f"video_{name}_keep"] = np.array(dec.remove_bad_images(frames), dtype=np.int64)
        fout[f"video_{name}_counts"] = np.array(
            [len(np.argwhere(cv2.absdiff(frames[k + 1], frames[k]) > 50)) for k in range(len(frames) - 1)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "frames_golden.npz"), **fout)

    for f in ("decode_golden.npz", "triangulate_golden.npz", "frames_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
