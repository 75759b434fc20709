"""Deterministic inputs for the post passes that follow the raster path (PassMotionBlur, PassLightShafts, TAA):
synthetic LDR / depth / motion planes that exercise the edge cases (velocities above max and below min, taps
leaving the frame, depth rejection of every tap, sun off screen / behind the camera, pass disabled)."""
import numpy as np

from leisure_software_renderer_b200 import capi


def planes(w, h, seed, max_motion=30.0):
    """LDR frame with smooth structure + noise, a depth plane with two layers and sky, motion with zero / small / huge regions."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    ldr = np.zeros((h, w, 4), dtype=np.uint8)
    base = (127.0 + 120.0 * np.sin(xx * 0.21 + seed) * np.cos(yy * 0.17)).astype(np.float32)
    for c in range(3):
        ldr[..., c] = np.clip(base + rng.integers(-40, 40, size=(h, w)) + 20 * c, 0, 255).astype(np.uint8)
    ldr[..., 3] = rng.integers(0, 256, size=(h, w)).astype(np.uint8)  # alpha is carried through by some branches
    depth = np.ones((h, w), dtype=np.float32)                          # 1.0 = sky
    depth[h // 4: 3 * h // 4, w // 5: 4 * w // 5] = 0.35
    depth[h // 3: h // 2, w // 3: w // 2] = 0.30 + rng.random((h // 2 - h // 3, w // 2 - w // 3), dtype=np.float32) * 0.1
    depth[0:3, :] = np.float32(1.7)                                     # outside [0,1]: the shafts clamp it
    depth[3:5, :] = np.float32(-0.2)
    motion = (rng.random((h, w, 2), dtype=np.float32) - 0.5) * np.float32(2.0 * max_motion)
    motion[:, : w // 4] *= np.float32(0.004)                            # below min_velocity_px
    motion[h // 2:, w // 2:] = 0.0
    motion[0, 0] = (np.float32(96.0), np.float32(-96.0))                # the raster path's clamp value
    return ldr, depth, motion


def blur_cases():
    P = capi.MotionBlurParams
    return {
        "default": (lambda: planes(96, 72, 1), P()),
        "strong_32": (lambda: planes(80, 50, 2, max_motion=90.0), P(samples=64, strength=2.5, max_velocity_px=40.0, min_velocity_px=0.0, depth_reject=0.02, dt=1 / 20)),
        "min_samples_tight_depth": (lambda: planes(33, 17, 3), P(samples=1, strength=0.7, max_velocity_px=0.5, depth_reject=0.0, dt=1e-6)),
        "disabled": (lambda: planes(40, 30, 4), P(enable=0)),
    }


def _cam(eye=(0.0, 3.0, -10.0), target=(0.0, 2.0, 0.0), aspect=4 / 3):
    from leisure_software_renderer_b200 import scenes
    return scenes.camera_viewproj(eye, target, (0.0, 1.0, 0.0), float(np.radians(60.0)), aspect, 0.1, 200.0), eye


def shafts_cases():
    def mk(sun_dir, with_depth=True, w=96, h=72, seed=5, **kw):
        vp, eye = _cam(aspect=w / h)
        p = capi.LightShaftsParams(cam_viewproj=vp, cam_pos=eye, sun_dir_ws=sun_dir, **kw)
        return (lambda: planes(w, h, seed)), p, with_depth
    n = lambda v: tuple(np.asarray(v, dtype=np.float32) / np.float32(np.linalg.norm(v)))
    return {
        "sun_in_view": mk(n((0.1, -0.25, -1.0))),
        "sun_in_view_no_depth": mk(n((-0.2, -0.1, -1.0)), with_depth=False, w=64, h=48, seed=6),
        "long_march": mk(n((0.3, -0.4, -1.0)), w=50, h=40, seed=7, steps=96, density=1.6, weight=0.4, decay=1.5),
        "few_steps": mk(n((0.0, -0.2, -1.0)), w=31, h=23, seed=8, steps=2, density=0.5, weight=2.0, decay=0.5),
        "sun_behind": mk(n((0.0, -0.3, 1.0))),
        "sun_off_screen": mk(n((-1.0, -0.2, -0.2))),
        "disabled": mk(n((0.1, -0.25, -1.0)), enable=0),
    }


def taa_frames(w=70, h=41, n=4, seed=11):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, size=(h, w, 4)).astype(np.uint8) for _ in range(n)]
