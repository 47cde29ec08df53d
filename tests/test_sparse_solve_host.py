"""The sparse static-order Newton solve of the form-1 generated kernels (specialize.gen_sparse_solve), compiled
for the HOST from the very text the generator writes into the cubin source: against numpy on the oracle's
Jacobians of real cells, and its fall-back contract (a multiplier over the threshold or a NaN: return false with the
matrix as assembled).  No GPU needed."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as orc  # noqa: E402
from pflotran_elm_interface_b200 import specialize, workloads as W  # noqa: E402

HARNESS = r"""
#include <cmath>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define SPEC_N %(n)d
#define SPEC_NC %(nc)d
#define SPEC_JS (SPEC_NC + 1)
#define JX(ci, cj) (((ci) * SPEC_JS + (cj)) * 32)
%(fn)s
extern "C" int pfrx_test_solve(const double *Jcore, const double *rhs, const double *c, double *out, double *Wback) {
  static double Wm[SPEC_NC * SPEC_JS * 32];
  std::memset(Wm, 0, sizeof(Wm));
  for (int i = 0; i < SPEC_NC; i++)
    for (int j = 0; j < SPEC_NC; j++) Wm[JX(i, j)] = Jcore[i * SPEC_NC + j];
  double res[SPEC_N], cc[SPEC_N];
  for (int i = 0; i < SPEC_N; i++) { res[i] = rhs[i]; cc[i] = c[i]; }
  const bool ok = spec_solve_sparse(Wm, res, cc);
  for (int i = 0; i < SPEC_N; i++) out[i] = res[i];
  for (int i = 0; i < SPEC_NC; i++)
    for (int j = 0; j < SPEC_NC; j++) Wback[i * SPEC_NC + j] = Wm[JX(i, j)];
  return ok ? 1 : 0;
}
"""


def _build(name, tmp_path):
    wl = W.by_name(name, ncell=24)
    g = specialize._Gen(wl.cfg)
    src = g.source()
    assert "#define SPEC_SPARSE_LU 1" in src
    m = re.search(r"__device__ __forceinline__ bool spec_solve_sparse\(.*?\n}\n", src, re.S)
    assert m, "generated source has no spec_solve_sparse"
    cpp = tmp_path / f"sparse_{name}.cpp"
    cpp.write_text(HARNESS % {"n": wl.cfg.ncomp, "nc": g.nc, "fn": m.group(0)})
    so = tmp_path / f"sparse_{name}.so"
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cpp)])
    lib = C.CDLL(str(so))
    dp = C.POINTER(C.c_double)
    lib.pfrx_test_solve.argtypes = [dp, dp, dp, dp, dp]
    return wl, g, lib


def _call(lib, g, n, Jc, rhs, c):
    out = np.zeros(n)
    back = np.zeros((g.nc, g.nc))
    dp = C.POINTER(C.c_double)
    ok = lib.pfrx_test_solve(np.ascontiguousarray(Jc).ctypes.data_as(dp), np.ascontiguousarray(rhs).ctypes.data_as(dp),
                             np.ascontiguousarray(c).ctypes.data_as(dp), out.ctypes.data_as(dp), back.ctypes.data_as(dp))
    return bool(ok), out, back


@pytest.mark.parametrize("name", ["c4fe", "c4s", "c4", "c3mr", "c6", "c7s", "c8"])
def test_generated_sparse_solve_matches_numpy(name, tmp_path):
    """the Newton update of the dense core from the generated straight-line elimination equals numpy's (RSolve's row
    scaling does not change the solution; in the log formulation the columns are scaled by c_j, reaction.F90:5493)"""
    wl, g, lib = _build(name, tmp_path)
    core = g.coupled
    n = wl.cfg.ncomp
    assert g.sparse_muladds < g.nc ** 3 / 3           # far fewer multiply-adds than a dense LU
    worst = 0.0
    for cell in range(24):
        st = wl.state.copy()
        e, R, J, _ = orc.girt_residual(wl.cfg, st, cell, wl.tran_dt)
        assert e == 0
        c = np.concatenate([st["pri_molal"][:, cell], st["immobile"][:, cell]])
        Jc = J[np.ix_(core, core)]
        ok, out, _ = _call(lib, g, n, Jc, R, c)
        assert ok, (name, cell)
        A = Jc * c[core][None, :] if wl.cfg.c.use_log_formulation else Jc
        x = out[core]
        # componentwise backward error (Oettli-Prager): the computed update solves a system whose entries differ
        # from A and r by this relative amount -- the criterion for a direct solver; the forward error against
        # numpy's LAPACK solution is that times the condition number (1e5-1e6 for the Hanford core)
        back = (np.abs(A @ x - R[core]) / (np.abs(A) @ np.abs(x) + np.abs(R[core]) + 1e-300)).max()
        worst = max(worst, back)
        want = np.linalg.solve(A, R[core])
        assert np.abs(x - want).max() <= 1.0e-8 * max(np.abs(want).max(), 1e-300), (name, cell)
        others = [i for i in range(n) if i not in core]
        assert np.array_equal(out[others], R[others])  # species outside the core are not this routine's business
    assert worst < 1.0e-13, (name, worst)


def test_generated_sparse_solve_falls_back_with_the_matrix_intact(tmp_path):
    """a multiplier over the threshold (here: a pivot 1e-9 of the entries below it) or a NaN makes the routine
    return false, and the matrix it leaves behind is the assembled one -- the dense pivoting LU runs on it next"""
    wl, g, lib = _build("c4fe", tmp_path)
    core, n = g.coupled, wl.cfg.ncomp
    st = wl.state.copy()
    _, R, J, _ = orc.girt_residual(wl.cfg, st, 3, wl.tran_dt)
    c = np.concatenate([st["pri_molal"][:, 3], st["immobile"][:, 3]])
    Jc = J[np.ix_(core, core)].copy()
    k = g.sparse_order[0]
    col = [i for i in range(g.nc) if i != k and Jc[i, k] != 0.0]
    assert col, "the first pivot has entries below it"
    bad = Jc.copy()
    bad[k, k] = 1.0e-9 * np.abs(bad[col, k]).max() / 64.0
    ok, _, back = _call(lib, g, n, bad, R, c)
    assert not ok
    assert np.array_equal(back, bad)
    nan = Jc.copy()
    nan[k, k] = np.nan
    ok, _, back = _call(lib, g, n, nan, R, c)
    assert not ok
    assert np.array_equal(np.isnan(back), np.isnan(nan)) and np.array_equal(np.nan_to_num(back), np.nan_to_num(nan))
