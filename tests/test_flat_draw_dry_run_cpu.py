"""Dry run, on the CPU, of the HOST logic of tests/test_zz_gpu_flat_draw.py: its bodies are executed against a stand-in context
whose flat-draw calls are served by the g++ build of the device functions (tests/cpp/flat_draw_emul.cpp) and whose light selection
is served by the restatement.  This checks the TEST CODE -- argument plumbing, array shapes, gates -- not the kernels; the product is
not involved (no libshsb.so compute call)."""
import numpy as np
import pytest

import test_flat_draw_cpu as fc
import test_zz_gpu_flat_draw as zf
from leisure_software_renderer_b200 import capi
from leisure_software_renderer_b200.renderer import Context
from oracle.bindings import SceneCull


class StandIn:
    def __init__(self):
        self.rts, self.meshes, self.emul = {}, {}, fc.emul()

    def rt_create(self, kind, w, h, zn=0.1, zf=1000.0):
        k = len(self.rts) + 1
        self.rts[k] = {capi.PLANE_COLOR: np.zeros((h, w, 4), np.uint8), capi.PLANE_DEPTH: np.zeros((h, w), np.float32)}
        return k

    def rt_destroy(self, rt):
        pass

    def rt_upload(self, rt, plane, array):
        self.rts[rt][plane][...] = np.asarray(array).reshape(self.rts[rt][plane].shape)

    def rt_download(self, rt, plane=capi.PLANE_COLOR):
        return self.rts[rt][plane].copy()

    def mesh_upload(self, positions, normals=None, uvs=None, indices=None):
        self.meshes[len(self.meshes) + 1] = (np.asarray(positions, np.float32).reshape(-1, 3), np.asarray(indices, np.uint32).reshape(-1))
        return len(self.meshes)

    _flat_draws = staticmethod(Context._flat_draws)

    def _run(self, mode, draws, view_proj, camera, light_dir, lights, canvas, depth):
        arr = self._flat_draws(draws)                      # through the ctypes structs the product receives
        table, verts, idx, fi, bv = [], [], [], 0, 0
        for k in sorted(self.meshes):
            v, i = self.meshes[k]
            table.append([fi, len(i), bv]); verts.append(v); idx.append(i)
            fi += len(i); bv += len(v)
        n = len(draws)
        sc = {"draw_mesh": np.array([arr[i].mesh - 1 for i in range(n)], np.uint32), "models": np.array([list(arr[i].model) for i in range(n)], np.float32).reshape(-1, 16),
              "base": np.array([list(arr[i].base_color) for i in range(n)], np.float32).reshape(-1, 3), "sel_counts": np.array([arr[i].selection_count for i in range(n)], np.uint32),
              "sel_idx": np.array([list(arr[i].selection) for i in range(n)], np.uint32).reshape(-1, 8), "mesh_table": np.array(table, np.uint32), "vertices": np.concatenate(verts),
              "indices": np.concatenate(idx), "view_proj": view_proj, "camera": camera, "light_dir": light_dir if light_dir is not None else np.zeros(3, np.float32),
              "lights": lights if lights is not None else np.zeros(0, capi.LIGHT_PROPS_DTYPE), "W": self.rts[canvas][capi.PLANE_COLOR].shape[1],
              "H": self.rts[canvas][capi.PLANE_COLOR].shape[0], "canvas": self.rts[canvas][capi.PLANE_COLOR], "depth": self.rts[depth][capi.PLANE_DEPTH]}
        if n:
            self.rts[canvas][capi.PLANE_COLOR], self.rts[depth][capi.PLANE_DEPTH] = self.emul.run(sc, mode)

    def flat_draw_blinn_phong(self, draws, view_proj, camera_pos, light_dir_ws, canvas_ldr, depth):
        self._run(0, draws, view_proj, camera_pos, light_dir_ws, None, canvas_ldr, depth)

    def flat_draw_multi_light(self, draws, view_proj, camera_pos, lights, canvas_ldr, depth):
        self._run(1, draws, view_proj, camera_pos, None, lights, canvas_ldr, depth)

    def collect_object_lights(self, aabbs, visible, records, mode):
        return SceneCull("port").collect_object_lights(aabbs, visible, records, mode)

    class _Lib:
        @staticmethod
        def shsb_mesh_destroy(h, m):
            return 0

    lib, h = _Lib(), None


@pytest.fixture(scope="module")
def standin():
    return StandIn()


@pytest.mark.parametrize("seed", [0, 1, 6, 7])
def test_dry_fuzz(standin, seed):
    zf.test_fuzz_flat_draws_equal_the_oracle.__wrapped__(standin, seed) if hasattr(zf.test_fuzz_flat_draws_equal_the_oracle, "__wrapped__") else zf.test_fuzz_flat_draws_equal_the_oracle(standin, seed)


def test_dry_split_batches(standin):
    zf.test_batches_compose_like_the_serial_loop(standin, 2)


def test_dry_selection_chain(standin):
    zf.test_selection_chain_feeds_the_draw(standin)


def test_dry_demo_sized_frame(standin):
    zf.test_demo_sized_frame_and_work_list_growth(standin)
