"""Seeded random inputs shared by the CPU fuzz suite (oracle vs compiled reference, bit-exact) and the GPU fuzz suite (CUDA vs
oracle).  Scenes themselves come from leisure_software_renderer_b200.scenes.scene_fuzz."""
import numpy as np

import post_cases
from leisure_software_renderer_b200 import capi, scenes


def motion_pair(seed, zero_normals=True):
    """Two consecutive random frames: every second item moves (sometimes far beyond the 96-px velocity clamp), the camera
    moves in 60 % of the cases; the second frame's prev_viewproj is the first frame's viewproj."""
    rng = np.random.default_rng(7000 + seed)
    prev = scenes.scene_fuzz(seed, zero_normals=zero_normals)
    prev.fp.motion_vectors_enable = 1
    kw = {}
    if rng.random() < 0.6:
        kw = {"cam_pos": tuple(float(a) + float(d) for a, d in zip(prev.cam_pos, rng.normal(0, 0.3, 3))), "cam_target": prev.cam_target}
    cur = prev.moved(dpos=tuple(float(v) for v in rng.normal(0, 1.0 if seed % 5 else 8.0, 3)), drot=tuple(float(v) for v in rng.normal(0, 0.4, 3)), **kw)
    return prev, cur


def post_inputs(seed):
    """Random planes + PassMotionBlur / PassLightShafts parameters, also outside their sane ranges (negative sample / step
    counts, negative strengths, dt = 0): returns (ldr, depth, motion, blur params, shafts params, shafts use depth)."""
    rng = np.random.default_rng(9000 + seed)
    w, h = int(rng.integers(3, 120)), int(rng.integers(3, 90))
    ldr, depth, motion = post_cases.planes(w, h, seed, max_motion=float(rng.choice([0.3, 8.0, 60.0, 300.0])))
    p = capi.MotionBlurParams(samples=int(rng.integers(-2, 70)), strength=float(rng.uniform(-0.5, 3.0)), max_velocity_px=float(rng.uniform(-1.0, 64.0)),
                              min_velocity_px=float(rng.uniform(0.0, 2.0)), depth_reject=float(rng.uniform(-0.05, 0.5)),
                              dt=float(rng.choice([0.0, 1e-6, 1 / 144, 1 / 30, 0.5])))
    eye = tuple(float(v) for v in rng.uniform(-5, 5, 3))
    tgt = tuple(float(v) for v in rng.uniform(-2, 2, 3))
    vp = scenes.camera_viewproj(eye, tgt, (0.0, 1.0, 0.0), float(np.radians(rng.uniform(30, 100))), w / h, 0.1, 200.0)
    sun = rng.normal(0, 1, 3)
    if seed % 3 == 0:   # the sun direction points away from where the sun is: aim it so that the sun projects into the view and the march runs
        sun = (np.asarray(tgt) - np.asarray(eye)) * -1.0 + rng.normal(0, 0.4, 3)
        sun = -sun
    q = capi.LightShaftsParams(cam_viewproj=vp, cam_pos=eye, sun_dir_ws=tuple(float(v) for v in sun / np.linalg.norm(sun)), steps=int(rng.integers(-1, 120)),
                               density=float(rng.uniform(-0.2, 2.0)), weight=float(rng.uniform(0.0, 2.5)), decay=float(rng.uniform(0.3, 1.6)))
    return ldr, depth, motion, p, q, bool(rng.random() < 0.8)


def light_bins(seed):
    """Random light set + camera + bin-builder arguments: returns (records, [(name, LightCullDesc, range_min, range_max)]) over
    the plain tiled builder's arguments and all four modes of shsb_light_cull_ex.  Viewports that are not tile multiples, tile
    sizes 8 / 16 / 20 / 32, caps from 1 to 128, ranges from centimetres to beyond the far plane, lights behind the camera, the
    sqrt(3)-inflated Jolt sphere bounds for a third of the sets, depth ranges that are thin, inverted-free and partly outside [0, 1]."""
    rng = np.random.default_rng(11000 + seed)
    w, h = int(rng.integers(20, 300)), int(rng.integers(16, 200))
    ts = int(rng.choice([8, 16, 20, 32]))
    cap = int(rng.choice([1, 4, 32, 128]))
    n = int(rng.integers(1, 260))
    n_spot = int(n * rng.uniform(0, 0.5))
    ext = float(rng.choice([4.0, 12.0, 40.0]))
    rl = float(rng.choice([0.05, 0.5, 2.0]))
    lights = scenes.make_lights(max(1, n - n_spot), n_spot, (-ext, -2.0, -ext), (ext, 4.0, ext), seed=seed, range_lo=rl,
                                range_hi=rl * float(rng.choice([2.0, 10.0, 200.0])), jolt_bounds=bool(seed % 3 == 0))
    eye = tuple(float(v) for v in rng.uniform(-ext, ext, 3) * np.array([1, 0.3, 1]))
    tgt = tuple(float(v) for v in rng.uniform(-ext / 2, ext / 2, 3))
    zn, zf = float(rng.choice([0.05, 0.1, 1.0])), float(rng.choice([15.0, 100.0, 1000.0]))
    vp = scenes.camera_viewproj(eye, tgt, (0.0, 1.0, 0.0), float(np.radians(rng.uniform(30, 110))), w / h, zn, zf)
    tiles = ((w + ts - 1) // ts) * ((h + ts - 1) // ts)
    a = rng.uniform(-0.1, 1.0, tiles).astype(np.float32)
    b = (a + rng.choice([0.0, 1e-6, 0.02, 0.5], tiles) * rng.uniform(0, 1, tiles)).astype(np.float32)
    va = (zn + (zf - zn) * np.clip(a, 0, 1) ** 2).astype(np.float32)
    vb = (va + rng.choice([0.0, 1e-4, 0.5, 30.0], tiles).astype(np.float32)).astype(np.float32)
    mk = lambda mode, **kw: capi.LightCullDesc(vp, w, h, mode, ts, cap, z_near=zn, z_far=zf, **kw)
    descs = [("tiled", mk(capi.LIGHT_CULL_TILED), None, None),
             ("depth01", mk(capi.LIGHT_CULL_TILED_DEPTH01), a, b),
             ("view_depth", mk(capi.LIGHT_CULL_TILED_VIEW_DEPTH), va, vb),
             ("clustered", mk(capi.LIGHT_CULL_CLUSTERED, depth_slices=int(rng.choice([1, 3, 16]))), None, None)]
    return lights, descs


def legacy_draws(seed):
    """Random inputs of the legacy tile-job rasterizer (BASELINE configs[0] as shipped): canvas size, job-tile size, one to three
    objects (Suzanne or a triangle soup with slivers and frustum-crossing triangles, expanded to per-corner streams like the demo's
    ModelGeometry), the demo's camera / world-matrix construction with random parameters -- often with geometry behind or
    around the camera, which this rasterizer projects through instead of clipping.
    Returns (W, H, tile_w, tile_h, camera_pos, light_dir, [(positions, normals, model_args, color)], view_args)."""
    rng = np.random.default_rng(13000 + seed)
    W, H = int(rng.integers(9, 200)), int(rng.integers(7, 150))
    tile_w, tile_h = (80, 80) if seed % 3 == 0 else (int(rng.integers(1, 90)), int(rng.integers(1, 90)))
    cam = tuple(float(v) for v in rng.uniform(-6, 6, 3))
    if seed % 4 == 0:
        cam = (float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1)))   # inside the objects
    h_ang, v_ang = float(rng.uniform(-180, 180)), float(rng.uniform(-70, 70))
    if seed % 2 == 0:  # look roughly at the origin
        d = -np.asarray(cam) / max(1e-3, np.linalg.norm(cam))
        h_ang = float(np.degrees(np.arctan2(d[0], d[2])) + rng.normal(0, 8))
        v_ang = float(np.degrees(np.arcsin(np.clip(d[1], -1, 1))) + rng.normal(0, 5))
    light = rng.normal(0, 1, 3)
    light = tuple(float(v) for v in light / np.linalg.norm(light))
    objs = []
    for k in range(int(rng.integers(1, 4))):
        if rng.random() < 0.5:
            m = scenes.load_suzanne()
            pos, nrm = m["positions"][m["indices"]], m["normals"][m["indices"]]
        else:
            m = scenes.make_triangle_soup(int(rng.integers(4, 80)), seed * 11 + k, extent=2.0, indexed=False, zero_normals=False)
            pos, nrm = m["positions"], m["normals"]
        scl = float(rng.uniform(0.3, 3.0))
        model_args = (tuple(float(v) for v in rng.uniform(-2, 2, 3)), (scl, scl * float(rng.choice([1.0, 1.0, 0.4])), scl), float(rng.uniform(-360, 360)))
        objs.append((np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(nrm, np.float32), model_args, tuple(int(v) for v in rng.integers(0, 256, 3)) + (255,)))
    return W, H, tile_w, tile_h, cam, light, objs, (h_ang, v_ang)


def _mat_look_at_lh(eye, center, up=(0.0, 1.0, 0.0)):
    eye, center, up = (np.asarray(v, np.float64) for v in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(up, f)
    s /= np.linalg.norm(s)
    u = np.cross(f, s)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, f
    m[0, 3], m[1, 3], m[2, 3] = -s @ eye, -u @ eye, -f @ eye
    return m


def legacy2_scene(seed):
    """Random inputs of the legacy soft-shadow demo: a floor quad grid + one to three objects (Suzanne or a
    soup), an orthographic light (z in 0..1 like the demo's ortho_lh_zo), a perspective camera that often clips the floor at the near
    plane, an optional random texture, random canvas / shadow-map / job-tile sizes.  All matrices are INPUTS of the path (float32,
    column-major when flattened): they only have to be the same for both sides.
    Returns dict(W, H, sm, tile, objs=[(pos, nrm, uv, model, color, use_tex)], view, proj, light_vp, light_dir, cam, texture)."""
    rng = np.random.default_rng(17000 + seed)
    W, H = int(rng.integers(24, 150)), int(rng.integers(20, 110))
    sm = int(rng.choice([24, 64, 150]))
    tile = (160, 160) if seed % 3 == 0 else (int(rng.integers(5, 90)), int(rng.integers(5, 90)))
    cam = np.array([rng.uniform(-6, 6), rng.uniform(0.3, 5), rng.uniform(-9, -2)])
    view = _mat_look_at_lh(cam, rng.uniform(-1, 1, 3) + np.array([0, 0.5, 0]))
    fov, asp, zn, zf = np.radians(rng.uniform(40, 90)), W / H, 0.1, 200.0
    t = np.tan(fov / 2)
    proj = np.zeros((4, 4))
    proj[0, 0], proj[1, 1], proj[2, 2], proj[3, 2], proj[2, 3] = 1 / (asp * t), 1 / t, zf / (zf - zn), 1.0, -zn * zf / (zf - zn)   # LH, z in 0..1
    ldir = rng.normal(0, 1, 3) * np.array([1, 0.3, 1]) - np.array([0, 0.9, 0])
    ldir /= np.linalg.norm(ldir)
    lview = _mat_look_at_lh(-ldir * 30.0, np.zeros(3))
    ext, lzn, lzf = float(rng.uniform(6, 14)), 0.1, 80.0
    lproj = np.eye(4)
    lproj[0, 0], lproj[1, 1], lproj[2, 2], lproj[2, 3] = 1 / ext, 1 / ext, 1 / (lzf - lzn), -lzn / (lzf - lzn)
    f32 = lambda m: np.ascontiguousarray(np.asarray(m, np.float32).T).reshape(16)     # row-major math -> column-major floats
    objs = []
    g = scenes.make_grid_plane(float(rng.uniform(8, 30)), int(rng.integers(1, 5)))
    objs.append((g["positions"][g["indices"]], g["normals"][g["indices"]], g["uvs"][g["indices"]] / 8.0, np.eye(4), tuple(int(v) for v in rng.integers(40, 256, 3)) + (255,), bool(rng.random() < 0.5)))
    for k in range(int(rng.integers(1, 4))):
        if rng.random() < 0.5:
            m = scenes.load_suzanne()
            pos, nrm, uv = m["positions"][m["indices"]], m["normals"][m["indices"]], m["uvs"][m["indices"]]
        else:
            m = scenes.make_triangle_soup(int(rng.integers(4, 60)), seed * 13 + k, extent=1.5, indexed=False, zero_normals=False)
            pos, nrm, uv = m["positions"], m["normals"], m["uvs"]
        model = np.eye(4)
        a, sc = rng.uniform(0, 6.28), rng.uniform(0.4, 2.0)
        model[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]) * sc * np.array([1, rng.choice([1.0, 0.5]), 1])
        model[:3, 3] = rng.uniform(-3, 3, 3) * np.array([1, 0.2, 1]) + np.array([0, 1.0, 0])
        objs.append((pos, nrm, uv, model, tuple(int(v) for v in rng.integers(0, 256, 3)) + (255,), bool(rng.random() < 0.4)))
    tw, th = int(rng.integers(1, 24)), int(rng.integers(1, 24))
    return {"W": W, "H": H, "sm": sm, "tile": tile, "objs": objs, "view": view, "proj": proj, "light_vp": f32(lproj @ lview), "light_dir": tuple(float(v) for v in ldir),
            "cam": tuple(float(v) for v in cam), "texture": rng.integers(0, 256, (th, tw, 4), dtype=np.uint8), "f32": f32}


def legacy3_scene(seed):
    """Random inputs of the legacy PBR / IBL demo: the L2 scene plus a previous-frame camera (motion vectors),
    per-object metallic / roughness / ao (also out of range), a random irradiance cube and a prefiltered-specular mip chain
    (HDR values; the library's cube-map sampling only reads them), IBL intensities.  Returns the L2 dict with extra keys."""
    sc = legacy2_scene(seed)
    rng = np.random.default_rng(23000 + seed)
    cam = np.asarray(sc["cam"])
    sc["prev_view"] = _mat_look_at_lh(cam + rng.normal(0, 0.15 if seed % 4 else 1.5, 3), rng.uniform(-1, 1, 3) + np.array([0, 0.5, 0])) if seed % 5 else sc["view"]
    sc["pbr"] = [(float(rng.uniform(-0.2, 1.2)), float(rng.uniform(-0.1, 1.3)), float(rng.uniform(-0.2, 1.2))) for _ in sc["objs"]]
    if seed % 7 == 6:
        sc["irradiance"], sc["prefiltered"] = None, None
    else:
        n_irr = int(rng.choice([1, 2, 8, 16]))
        sc["irradiance"] = rng.uniform(0, 2.5, (6, n_irr, n_irr, 3)).astype(np.float32)
        base = int(rng.choice([1, 4, 16, 32]))
        n_mips = int(rng.integers(1, 7))
        sc["prefiltered"] = [rng.uniform(0, 4.0, (6, max(1, base >> m), max(1, base >> m), 3)).astype(np.float32) for m in range(n_mips)]
    sc["ibl_k"] = (float(rng.uniform(-0.2, 1.4)), float(rng.uniform(-0.2, 1.4)), float(rng.uniform(0, 1.2)))
    return sc


def scene_cull(seed):
    """Random inputs of the scene-level steps upstream of draw submission: object AABBs around / inside / behind
    the camera frustum (thin slabs, huge boxes that contain the camera, point-sized boxes, boxes touching a plane), a camera, a
    light set and a visible-light list with repeats, gaps and out-of-range entries; ties in distance come from lights sharing a position.
    Returns dict(aabbs (n, 6), view_proj, view, lights, visible (lights), visible_objects, w, h, ts, zn, zf)."""
    rng = np.random.default_rng(29000 + seed)
    n = int(rng.integers(1, 400))
    ext = float(rng.choice([3.0, 15.0, 60.0]))
    c = rng.uniform(-ext, ext, (n, 3)) * np.array([1, 0.4, 1])
    half = np.abs(rng.normal(0, 1, (n, 3))) * rng.choice([0.01, 0.5, 3.0, 30.0], (n, 1)) * rng.choice([1.0, 1.0, 0.02], (n, 3))
    half[rng.random(n) < 0.05] = 0.0
    aabbs = np.concatenate([c - half, c + half], axis=1).astype(np.float32)
    eye = tuple(float(v) for v in rng.uniform(-ext, ext, 3) * np.array([1, 0.3, 1]))
    tgt = tuple(float(v) for v in rng.uniform(-ext / 2, ext / 2, 3))
    zn, zf = float(rng.choice([0.05, 0.1, 1.0])), float(rng.choice([15.0, 100.0, 1000.0]))
    vp = scenes.camera_viewproj(eye, tgt, (0.0, 1.0, 0.0), float(np.radians(rng.uniform(30, 110))), float(rng.uniform(0.6, 2.2)), zn, zf)
    nl = int(rng.integers(1, 200))
    lights = scenes.make_lights(max(1, nl - nl // 4), nl // 4, (-ext, -2.0, -ext), (ext, 4.0, ext), seed=seed + 7, range_lo=0.2,
                                range_hi=float(rng.choice([1.0, 8.0, 80.0])), jolt_bounds=bool(seed % 2))
    if len(lights) > 12:                                  # equal positions -> equal squared distances (the strict-less rule decides)
        for k in range(0, len(lights) - 1, 5):
            lights["position_range"][k + 1, :3] = lights["position_range"][k, :3]
    nv = int(rng.integers(0, 3 * len(lights)))
    visible = rng.integers(0, len(lights) + 3, nv).astype(np.uint32) if seed % 3 else np.arange(len(lights), dtype=np.uint32)
    view = np.ascontiguousarray(scenes.look_at_lh(eye, tgt, (0.0, 1.0, 0.0)).astype(np.float32).T).reshape(16)
    w, h, ts = int(rng.integers(20, 400)), int(rng.integers(16, 260)), int(rng.choice([8, 16, 20, 32]))
    nvo = int(rng.integers(0, 2 * n))
    visible_objects = rng.integers(0, n + 2, nvo).astype(np.uint32) if seed % 4 else np.arange(n, dtype=np.uint32)
    return {"aabbs": aabbs, "view_proj": vp, "lights": lights, "visible": visible, "view": view, "w": w, "h": h, "ts": ts, "zn": zn, "zf": zf,
            "visible_objects": visible_objects}


def occlusion_scene(seed):
    """Random input of the software-occlusion pass (geometry/culling_software.hpp): a camera looking down a corridor of boxes and
    spheres-as-boxes of very different sizes -- big walls near the camera that hide what is behind them, small props, objects
    behind / around the camera, objects with no occluder mesh, a mesh with a dangling index pair (indices.size() % 3 != 0), stale
    entries in the frustum-visible list (indices >= n_objects) -- and occlusion buffers from 8 x 6 to 320 x 180.  Occluder meshes
    are DebugMesh-like: a unit box (12 triangles), an octahedron (8), a thin quad (2), each drawn with the object's model matrix."""
    rng = np.random.default_rng(19000 + seed)
    box_v = np.array([[x, y, z] for z in (-.5, .5) for y in (-.5, .5) for x in (-.5, .5)], np.float32)
    box_i = np.array([0, 1, 3, 0, 3, 2, 4, 6, 7, 4, 7, 5, 0, 4, 5, 0, 5, 1, 2, 3, 7, 2, 7, 6, 0, 2, 6, 0, 6, 4, 1, 5, 7, 1, 7, 3], np.uint32)
    oct_v = np.array([[.5, 0, 0], [-.5, 0, 0], [0, .5, 0], [0, -.5, 0], [0, 0, .5], [0, 0, -.5]], np.float32)
    oct_i = np.array([0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5], np.uint32)
    quad_v = np.array([[-.5, -.5, 0], [.5, -.5, 0], [.5, .5, 0], [-.5, .5, 0]], np.float32)
    quad_i = np.array([0, 1, 2, 0, 2, 3, 1, 2], np.uint32)            # 8 indices: the trailing pair is ignored (i + 2 < size)
    vertices = np.concatenate([box_v, oct_v, quad_v])
    indices = np.concatenate([box_i, oct_i, quad_i])
    mesh_table = np.array([[0, 36, 0], [36, 24, 8], [60, 8, 14]], np.uint32)
    n = int(rng.integers(1, 120))
    ext = float(rng.choice([6.0, 20.0, 60.0]))
    aabbs, models, omesh = [], [], []
    for i in range(n):
        c = rng.uniform(-ext, ext, 3) * np.array([1.0, 0.4, 1.0])
        big = rng.random() < 0.15
        half = rng.uniform(0.1, 1.5, 3) * (float(rng.uniform(3, 10)) if big else 1.0)
        m = int(rng.choice([0, 0, 1, 2, 0xFFFFFFFF])) if not big else 0
        model = np.eye(4, dtype=np.float32)
        model[0, 0], model[1, 1], model[2, 2] = 2 * half          # column-major storage below: scale then translate
        model[:3, 3] = c
        aabbs.append(np.concatenate([c - half, c + half]))
        models.append(model.T.reshape(16))                          # glm column-major
        omesh.append(m)
    aabbs = np.array(aabbs, np.float32)
    eye = tuple(float(v) for v in rng.uniform(-ext, ext, 3) * np.array([1.0, 0.3, 1.0]))
    tgt = tuple(float(v) for v in rng.uniform(-ext / 3, ext / 3, 3))
    w, h = [(8, 6), (64, 36), (160, 90), (320, 180), (97, 53)][int(rng.integers(0, 5))]
    zn, zf = float(rng.choice([0.05, 0.1, 1.0])), float(rng.choice([50.0, 300.0, 1000.0]))
    view_proj = scenes.camera_viewproj(eye, tgt, (0.0, 1.0, 0.0), float(np.radians(rng.uniform(40, 90))), w / h, zn, zf)
    view = np.ascontiguousarray(scenes.look_at_lh(eye, tgt, (0.0, 1.0, 0.0)).astype(np.float32).T).reshape(16)
    visible = rng.permutation(n)[: int(rng.integers(0, n + 1))].astype(np.uint32)
    if seed % 4 == 0 and len(visible):
        visible = np.concatenate([visible, np.array([n + 3, 0xFFFFFFF0], np.uint32)])      # stale indices
        rng.shuffle(visible)
    return {"aabbs": aabbs, "visible": visible, "object_mesh": np.array(omesh, np.uint32), "models": np.array(models, np.float32), "mesh_table": mesh_table,
            "vertices": vertices, "indices": indices, "view": view, "view_proj": view_proj, "occ_w": w, "occ_h": h, "eps": float(rng.choice([1e-4, 0.0, 1e-2]))}


def _uv_sphere(n_lon=14, n_lat=9):
    """A DebugMesh-like unit sphere (radius 0.5): the tessellated shapes the demos draw (spheres, capsules) are of this kind."""
    v = [[0.0, 0.5, 0.0]]
    for j in range(1, n_lat):
        th = np.pi * j / n_lat
        for i in range(n_lon):
            ph = 2 * np.pi * i / n_lon
            v.append([0.5 * np.sin(th) * np.cos(ph), 0.5 * np.cos(th), 0.5 * np.sin(th) * np.sin(ph)])
    v.append([0.0, -0.5, 0.0])
    idx = []
    ring = lambda j, i: 1 + (j - 1) * n_lon + (i % n_lon)
    for i in range(n_lon):
        idx += [0, ring(1, i + 1), ring(1, i)]
        idx += [len(v) - 1, ring(n_lat - 1, i), ring(n_lat - 1, i + 1)]
    for j in range(1, n_lat - 1):
        for i in range(n_lon):
            idx += [ring(j, i), ring(j, i + 1), ring(j + 1, i + 1), ring(j, i), ring(j + 1, i + 1), ring(j + 1, i)]
    return np.array(v, np.float32), np.array(idx, np.uint32)


def _rotation(rng):
    a = rng.normal(size=3)
    a /= max(np.linalg.norm(a), 1e-6)
    t = rng.uniform(0, 2 * np.pi)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(t) * K + (1 - np.cos(t)) * (K @ K)


def flat_draw_lights(rng, n, ext):
    """n lights of the four local light models (lighting/light_runtime.hpp:291-535) with properties around and beyond the
    clamps their sample() functions apply: ranges from smaller than the scene to larger, every attenuation model, powers
    below / at / above 1, cut-offs, negative colours and intensities, zero direction / axis vectors (the fallbacks), cone angles
    outside [0.02, pi/2 - 0.02] and outer < inner."""
    li = np.zeros(n, capi.LIGHT_PROPS_DTYPE)
    for i in range(n):
        l = li[i]
        l["light_type"] = int(rng.integers(1, 5))
        l["color"] = rng.uniform(0.1, 1.0, 3) if rng.random() > 0.08 else rng.uniform(-0.5, 1.0, 3)
        l["intensity"] = float(rng.uniform(0.3, 6.0)) if rng.random() > 0.05 else -1.0
        l["position_ws"] = rng.uniform(-ext, ext, 3) * np.array([1.0, 0.35, 1.0]) + np.array([0.0, 0.3 * ext, 0.0])
        l["range"] = float(rng.choice([0.4, 1.0, 2.0, 4.0]) * ext * rng.uniform(0.3, 1.0))
        d = rng.normal(size=3)
        l["direction_ws"] = d / np.linalg.norm(d) * float(rng.choice([1.0, 1.0, 3.0])) if rng.random() > 0.06 else np.zeros(3)
        l["inner_angle_rad"] = float(rng.choice([0.0, 0.2, 0.5, 1.0, 1.7]))
        l["outer_angle_rad"] = float(rng.choice([0.1, 0.4, 0.8, 1.3, 2.0]))
        r = rng.normal(size=3)
        l["right_ws"] = r / np.linalg.norm(r) if rng.random() > 0.06 else np.zeros(3)
        # an up hint that is never parallel to the direction: the parallel case divides by zero inside right_from_forward and ends in a
        # float -> uint8 conversion of NaN, which C++ leaves undefined
        u = np.cross(l["direction_ws"] if np.any(l["direction_ws"]) else np.array([0.0, -1.0, 0.0]), rng.normal(size=3))
        l["up_ws"] = u / max(np.linalg.norm(u), 1e-6) + 0.2 * rng.normal(size=3)
        l["tube_half_length"] = float(rng.choice([0.02, 0.5, 1.5, 4.0]))
        l["rect_half_extents"] = rng.choice([0.01, 0.4, 1.0, 3.0], 2)
        l["tube_radius"] = 0.25
        l["attenuation_model"] = int(rng.integers(0, 3)) if rng.random() > 0.04 else 7
        l["attenuation_power"] = float(rng.choice([1.0, 1.0, 0.5, 2.0, 3.3, 0.0]))
        l["attenuation_bias"] = float(rng.choice([0.05, 0.0, 1.0]))
        l["attenuation_cutoff"] = float(rng.choice([0.0, 0.0, 0.02, 0.3]))
        l["flags"] = 1
    return li


def flat_draw_scene(seed, dangling=False, size=None):
    """Random batch of flat-shaded draws (sw_render/debug_draw.hpp:153-203; hello_light_types_culling_sw.cpp:366-422): boxes,
    octahedra, quads (with a trailing index pair that is ignored), tessellated spheres and a floor that covers the canvas, under
    rotated / non-uniformly scaled model matrices; objects behind and across the near plane (a triangle with one rejected vertex is
    skipped whole); duplicated draws with another colour (equal depths: the earlier draw stays); degenerate triangles; per-draw
    light selections of 0..8 entries with stale indices; a canvas that is not empty and a depth buffer that already holds a
    near block.  dangling=True adds a mesh whose indices point beyond the vertex array (skipped by the restatement and the device; the
    reference would read out of bounds, so those scenes are not fed to it)."""
    rng = np.random.default_rng(23000 + seed)
    box_v = np.array([[x, y, z] for z in (-.5, .5) for y in (-.5, .5) for x in (-.5, .5)], np.float32)
    box_i = np.array([0, 1, 3, 0, 3, 2, 4, 6, 7, 4, 7, 5, 0, 4, 5, 0, 5, 1, 2, 3, 7, 2, 7, 6, 0, 2, 6, 0, 6, 4, 1, 5, 7, 1, 7, 3], np.uint32)
    oct_v = np.array([[.5, 0, 0], [-.5, 0, 0], [0, .5, 0], [0, -.5, 0], [0, 0, .5], [0, 0, -.5]], np.float32)
    oct_i = np.array([0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5], np.uint32)
    quad_v = np.array([[-.5, 0, -.5], [.5, 0, -.5], [.5, 0, .5], [-.5, 0, .5]], np.float32)
    quad_i = np.array([0, 2, 1, 0, 3, 2, 1, 2], np.uint32)             # 8 indices: the trailing pair is ignored (i + 2 < size)
    sph_v, sph_i = _uv_sphere()
    deg_v = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [0, 1, 0]], np.float32)
    deg_i = np.array([0, 1, 2, 0, 0, 3, 0, 1, 3], np.uint32)           # collinear, repeated vertex, one proper triangle
    meshes = [(box_v, box_i), (oct_v, oct_i), (quad_v, quad_i), (sph_v, sph_i), (deg_v, deg_i)]
    if dangling:
        meshes.append((oct_v, np.array([0, 2, 4, 2, 1, 40, 1, 3, 4, 0xFFFFFFFF, 0, 4, 2, 0, 5], np.uint32)))  # LAST mesh: indices beyond the array
    vertices = np.concatenate([m[0] for m in meshes])
    indices = np.concatenate([m[1] for m in meshes])
    table, fi, bv = [], 0, 0
    for v, i in meshes:
        table.append([fi, len(i), bv])
        fi += len(i); bv += len(v)
    ext = float(rng.choice([5.0, 12.0, 30.0]))
    n = int(rng.integers(1, 50))
    draw_mesh, models, base = [], [], []
    for k in range(n):
        if k and rng.random() < 0.12:                                   # the same object again in another colour: every depth ties
            draw_mesh.append(draw_mesh[-1]); models.append(models[-1]); base.append(rng.uniform(0, 1, 3))
            continue
        floor = rng.random() < 0.08
        m = 2 if floor else int(rng.choice([0, 0, 1, 2, 3, 3, 4] + ([5] if dangling else [])))
        scl = np.array([4 * ext, 1.0, 4 * ext]) if floor else rng.uniform(0.2, 2.5, 3) * float(rng.choice([1.0, 1.0, 3.0]))
        R = np.eye(3) if floor else _rotation(rng)
        M = np.eye(4)
        M[:3, :3] = R @ np.diag(scl)
        M[:3, 3] = np.array([0.0, -0.2 * ext, 0.0]) if floor else rng.uniform(-ext, ext, 3) * np.array([1.0, 0.35, 1.0])
        draw_mesh.append(m); models.append(M.T.astype(np.float32).reshape(16)); base.append(rng.uniform(0, 1.1, 3))
    eye = rng.uniform(-ext, ext, 3) * np.array([1.0, 0.3, 1.0]) + np.array([0.0, 0.25 * ext, 0.0])
    tgt = rng.uniform(-ext / 3, ext / 3, 3)
    w, h = size or [(16, 12), (97, 53), (160, 90), (320, 180), (200, 150)][int(rng.integers(0, 5))]
    zn, zf = float(rng.choice([0.05, 0.1, 1.0])), float(rng.choice([50.0, 300.0]))
    view_proj = scenes.camera_viewproj(tuple(map(float, eye)), tuple(map(float, tgt)), (0.0, 1.0, 0.0), float(np.radians(rng.uniform(40, 90))), w / h, zn, zf)
    n_lights = int(rng.integers(0, 25))
    lights = flat_draw_lights(rng, n_lights, ext)
    sel_counts = rng.integers(0, 9, n).astype(np.uint32)
    sel_idx = rng.integers(0, n_lights + 2, (n, 8)).astype(np.uint32)  # n_lights, n_lights + 1: stale entries (skipped, :411)
    if seed % 7 == 0:
        sel_idx[0, 0] = 0xFFFFFFFF
    canvas = np.empty((h, w, 4), np.uint8)
    canvas[:] = np.array([int(rng.integers(0, 60)), int(rng.integers(0, 60)), int(rng.integers(0, 80)), 255], np.uint8)
    depth = np.ones((h, w), np.float32)
    if seed % 3 == 0:
        depth[h // 4: h // 2, w // 3: 2 * w // 3] = np.float32(rng.choice([0.0, 0.5, 0.97]))
    ld = rng.normal(size=3)
    return {"draw_mesh": np.array(draw_mesh, np.uint32), "models": np.array(models, np.float32), "base": np.array(base, np.float32), "sel_counts": sel_counts, "sel_idx": sel_idx,
            "mesh_table": np.array(table, np.uint32), "vertices": vertices, "indices": indices, "view_proj": view_proj, "camera": eye.astype(np.float32),
            "light_dir": (ld / np.linalg.norm(ld)).astype(np.float32) * np.float32(rng.choice([1.0, 2.5])), "lights": lights, "W": w, "H": h, "canvas": canvas, "depth": depth}
