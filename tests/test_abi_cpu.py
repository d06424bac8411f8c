"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/d2dx.h declares,
validates arguments without touching a GPU, and the host-side packing logic is right."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "d2dx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(d2dx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from d2d_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/d2dx.h but not exported by libd2dx.so"
    assert sorted(_lib.EXPORTED) == declared, "ctypes binding and header disagree"
    version = int(re.search(r"#define\s+D2DX_VERSION\s+(\d+)", open(os.path.join(ROOT, "include", "d2dx.h")).read()).group(1))
    assert _lib.lib.d2dx_version() == version and version >= 102


def test_struct_sizes_match_the_header():
    """ctypes mirrors of the ABI structs have the C layout (compiled with gcc from the header itself)."""
    import subprocess, tempfile
    from d2d_b200 import _lib
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "d2dx.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(d2dx_traj_table), sizeof(d2dx_dfff_gains), sizeof(d2dx_scenarios),
  sizeof(d2dx_rollout_out), sizeof(d2dx_formations), sizeof(d2dx_formation_out), sizeof(d2dx_colloc_problem),
  offsetof(d2dx_colloc_problem, obs), offsetof(d2dx_scenarios, pert_begin)); return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()
    got = [C.sizeof(_lib.TrajTable), C.sizeof(_lib.DfffGains), C.sizeof(_lib.Scenarios), C.sizeof(_lib.RolloutOut),
           C.sizeof(_lib.Formations), C.sizeof(_lib.FormationOut), C.sizeof(_lib.CollocProblem),
           _lib.CollocProblem.obs.offset, _lib.Scenarios.pert_begin.offset]
    assert [int(v) for v in out] == got


def test_argument_validation_without_gpu():
    from d2d_b200 import _lib
    lib = _lib.lib
    g = _lib.DfffGains()
    assert lib.d2dx_dfff_default_gains(C.byref(g)) == 0
    assert (g.q_pos, g.q_psi, g.r_phi, g.r_v) == (1.0, 0.1, 8.0, 1.0)
    assert abs(g.u_hi[0] - np.deg2rad(45)) < 1e-16 and g.u_lo[1] == 4.0 and g.u_hi[1] == 20.0
    assert lib.d2dx_dfff_default_gains(None) == 1
    p = _lib.CollocProblem(); p.n_ac, p.N, p.h, p.n_inst = 16, 500, 0.02, 96
    s = (C.c_int64 * 3)()
    assert lib.d2dx_colloc_sizes(C.byref(p), _lib.JAC_COMPACT, s) == 0
    assert list(s) == [40000, 48 * 499 + 96, 12 * 16 * 499 + 96]
    assert lib.d2dx_colloc_sizes(C.byref(p), _lib.JAC_OPTY_DENSE, s) == 0
    assert s[2] == 499 * 48 * 128 + 96
    assert lib.d2dx_colloc_scratch_size(C.byref(p), 3) > 0
    h = C.c_void_p()
    rc = lib.d2dx_create(0, C.byref(h))                              # no GPU here: must fail loudly, not fall back
    import torch
    if not torch.cuda.is_available():
        assert rc != 0 and len(lib.d2dx_last_error()) > 0
        with pytest.raises(RuntimeError):
            import d2d_b200
            d2d_b200.Engine()
    else:
        lib.d2dx_destroy(h)


def test_trajectory_packing():
    from d2d_b200 import _lib, trajectory as ddt, trajectory_factory as ddtf
    sq = ddtf.TrajSquare()
    two = ddtf.TrajTwoLines()
    circ = ddt.TrajectoryCircle(c=[1., 2.], r=-25., v=10., alpha0=0.3)
    p = ddt.pack([circ, sq, two])
    assert p.uniform_type == -1 and list(p.first_seg) == [0, 1, 5] and list(p.n_segs) == [1, 4, 2]
    assert p.traj_dur[0] == 0. and p.traj_dur[1] == 20. and abs(p.traj_dur[2] - 2 * np.sqrt(5000) / 10) < 1e-12
    np.testing.assert_array_equal(p.seg_end[1:5], [5., 10., 15., 20.])
    assert list(p.seg_type) == [_lib.SEG_CIRCLE] + [_lib.SEG_LINE] * 6
    np.testing.assert_array_equal(p.seg_par[:6, 0], [0., 1., 2., -25., 10. / -25., 0.3])
    np.testing.assert_array_equal(p.seg_par[0, 1:5], [0., 5., 10., 15.])         # steps reset to the previous end
    assert ddt.pack([circ, circ]).uniform_type == _lib.SEG_CIRCLE
    b = ddt.CircleBatch([0., 1.], [2., 3.], [30., 40.], [10., 12.], [0., 1.])
    pb = b.pack()
    assert pb.uniform_type == _lib.SEG_CIRCLE and pb.seg_par.shape == (_lib.SEG_NPAR, 2)
    np.testing.assert_array_equal(pb.seg_par[4], [10. / 30., 12. / 40.])
    ms = ddtf.TrajMinSnapDemo()
    pm = ddt.pack([ms])
    assert pm.uniform_type == _lib.SEG_POLY
    np.testing.assert_array_equal(pm.seg_par[1:9, 0], ms._polys[0].coefs[0])
    with pytest.raises(NotImplementedError):
        ddt.pack([ddt.CompositeTraj([sq, circ])])
    from oracle import d2d_oracle as orc
    np.testing.assert_array_equal(ms._polys[0].coefs, orc.traj_minsnap_demo()._polys[0].coefs)


def test_shard_bookkeeping():
    from d2d_b200.distributed import AircraftShard, shard_range
    assert [shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sum(b - a for a, b in (shard_range(10 ** 6, 8, r) for r in range(8))) == 10 ** 6
    n_ac, N = 4, 5
    inst = [(3 * a + k, 0, 1.) for a in range(n_ac) for k in range(3)] + [(3 * a + k, N - 1, 2.) for a in range(n_ac) for k in range(3)]
    seen_free, seen_con, seen_jac = [], [], []
    for r in range(2):
        s = AircraftShard(n_ac, N, inst, 2, r)
        assert s.n_own == 2 and len(s.inst_local) == 12 and all(0 <= v < 6 for v, _, _ in s.inst_local)
        seen_free += list(s.idx_free); seen_con += list(s.idx_con); seen_jac += list(s.idx_jac)
    assert sorted(seen_free) == list(range(5 * n_ac * N))
    assert sorted(seen_con) == list(range(3 * n_ac * (N - 1) + len(inst)))
    assert sorted(seen_jac) == list(range(12 * n_ac * (N - 1) + len(inst)))
