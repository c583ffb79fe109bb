"""CPU-side checks of the drop-in boundary: libdemcz_cuda.so loads, exports every symbol that
include/demcz.h declares, its POD structs match the ctypes mirror, and — with no CUDA device —
the product path fails loudly instead of falling back to anything on the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import demc_jl_b200 as DEMC
from demc_jl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = torch.cuda.is_available()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "demcz.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(demcz_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libdemcz_cuda.so does not export {n}"
    assert sorted(_lib.EXPORTS) == names


def test_version_and_struct_sizes():
    lib = _lib.load()
    assert lib.demcz_version() == 200
    # the library rejects a mismatching struct_size: proves both sides agree on sizeof(demcz_opts)
    o = _lib.Opts()
    o.struct_size = C.sizeof(_lib.Opts) - 4
    ctx = C.c_void_p()
    assert lib.demcz_create(C.byref(ctx), C.byref(o), 0) == -1
    assert b"struct_size" in lib.demcz_last_error(None)


def test_create_validates_arguments_before_touching_cuda():
    lib = _lib.load()
    bp = np.array([0, 2], dtype=np.int32)
    bi = np.array([0, 5], dtype=np.int32)          # column 5 outside d=2
    eps = np.ones(2)
    o = _lib.Opts()
    o.struct_size = C.sizeof(_lib.Opts)
    o.N, o.K, o.Ngeneration, o.Nblocks, o.d = 4, 10, 10, 1, 2
    o.block_ptr = bp.ctypes.data_as(C.POINTER(C.c_int32))
    o.block_idx = bi.ctypes.data_as(C.POINTER(C.c_int32))
    o.eps_scale = eps.ctypes.data_as(C.POINTER(C.c_double))
    o.rank, o.world, o.N_global = 0, 1, 4
    ctx = C.c_void_p()
    assert lib.demcz_create(C.byref(ctx), C.byref(o), 0) == -1
    assert b"outside" in lib.demcz_last_error(None)


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    mu = np.zeros(3)
    with pytest.raises(_lib.DemczError) as ei:
        DEMC.demcz_sample(DEMC.MvNormalLogpdf(mu, np.eye(3)), np.random.default_rng(0).normal(size=(30, 3)),
                          4, 10, 20, 1, [[0, 1, 2]], 1e-4 * np.ones(3), 2.38, verbose=False)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_closure_logobj_is_rejected():
    with pytest.raises(TypeError):
        DEMC.demcz_sample(lambda x: -np.sum(x * x), np.zeros((10, 2)), DEMC.demcopt(2))


def test_demcopt_defaults_match_reference():
    # src/DEMC.jl:41-43
    o = DEMC.demcopt(7)
    assert (o.N, o.K, o.Ngeneration, o.Nblocks) == (4, 10, 5000, 1)
    assert o.blockindex == [list(range(7))]
    np.testing.assert_array_equal(o.eps_scale, 1e-4 * np.ones(7))
    assert (o.γ, o.verbose, o.print_step, o.T0, o.TN) == (2.38, True, 100, 3, 1e-3)
    assert (o.autostop, o.autostop_every, o.autostop_Rhat) == ("Rhat", 1000, 1.05)
    o.γ = 2.0
    assert o.gamma == 2.0


def test_host_diagnostics_match_oracle():
    from oracle import demcz_oracle as orc
    rng = np.random.default_rng(0)
    chain = rng.normal(size=(6, 3, 41)) + rng.normal(size=(6, 3, 1))
    np.testing.assert_allclose(DEMC.Rhat_gelman(chain), orc.rhat_gelman(chain), rtol=1e-13)
    np.testing.assert_array_equal(DEMC.flatten_chain(chain), orc.flatten_chain(chain))
    b, c = DEMC.mean_cov_chain(chain)
    b2, c2 = orc.mean_cov_chain(chain)
    np.testing.assert_allclose(b, b2)
    np.testing.assert_allclose(c, c2)
    lo = rng.integers(0, 3, size=(6, 41)).astype(float)
    ar, rh = DEMC.convergence_check(chain, lo, verbose=False)
    np.testing.assert_array_equal(ar, orc.accept_ratio(lo))


def test_header_is_plain_c_and_struct_layout_matches_ctypes(tmp_path):
    """include/demcz.h must be consumable by a C compiler (the Julia ccall / cgo style binding sees a C ABI), and the
    ctypes mirror must agree with the compiler on every field offset of demcz_opts / demcz_run_info."""
    import subprocess
    fields_o = [n for n, _ in _lib.Opts._fields_]
    fields_i = [n for n, _ in _lib.RunInfo._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "demcz.h"', 'int main(void) {',
           '  printf("%zu %zu\\n", sizeof(demcz_opts), sizeof(demcz_run_info));']
    src += [f'  printf("o {f} %zu\\n", offsetof(demcz_opts, {f}));' for f in fields_o]
    src += [f'  printf("i {f} %zu\\n", offsetof(demcz_run_info, {f}));' for f in fields_i]
    src += ['  return (int)(DEMCZ_XCHG_DENSE_SLOTS + DEMCZ_SYNC_OVERLAP + DEMCZ_ERR_UNSUPPORTED) * 0;', '}']
    cfile = tmp_path / "abi.c"
    cfile.write_text("\n".join(src))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(cfile)],
                   check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    so, si = map(int, out[0].split())
    assert so == C.sizeof(_lib.Opts) and si == C.sizeof(_lib.RunInfo)
    for line in out[1:]:
        if not line:
            continue
        which, name, off = line.split()
        cls = _lib.Opts if which == "o" else _lib.RunInfo
        assert getattr(cls, name).offset == int(off), (which, name)


def test_create_rejects_unknown_exchange_layout():
    lib = _lib.load()
    bp = np.array([0, 2], dtype=np.int32)
    bi = np.array([0, 1], dtype=np.int32)
    eps = np.ones(2)
    o = _lib.Opts()
    o.struct_size = C.sizeof(_lib.Opts)
    o.N, o.K, o.Ngeneration, o.Nblocks, o.d = 4, 10, 10, 1, 2
    o.block_ptr = bp.ctypes.data_as(C.POINTER(C.c_int32))
    o.block_idx = bi.ctypes.data_as(C.POINTER(C.c_int32))
    o.eps_scale = eps.ctypes.data_as(C.POINTER(C.c_double))
    o.rank, o.world, o.N_global = 0, 1, 4
    o.exchange_layout = 9
    ctx = C.c_void_p()
    assert lib.demcz_create(C.byref(ctx), C.byref(o), 0) == -1
    assert b"exchange_layout" in lib.demcz_last_error(None)
    o.exchange_layout = 2                            # the removed `packed` staging-ring exchange
    assert lib.demcz_create(C.byref(ctx), C.byref(o), 0) == -1
    assert b"exchange_layout" in lib.demcz_last_error(None)
    for bad in ("bogus", "packed"):
        with pytest.raises(ValueError):
            DEMC.Sampler(DEMC.MvNormalLogpdf(np.zeros(2), np.eye(2)), 2, N=4, K=10, Ngeneration=10, blockindex=[[0, 1]],
                         eps_scale=np.ones(2), γ=2.38, exchange_layout=bad)


def test_extract_best_returns_first_maximum_of_the_history():
    rng = np.random.default_rng(3)
    chain = rng.normal(size=(5, 3, 7))
    lo = rng.normal(size=(5, 7))
    lo[2, 4] = lo[4, 1] = 10.0                       # a tie: column-major order picks generation 1 (chain 4) first
    mc = DEMC.MC(chain, lo, chain[:, :, -1], lo[:, -1])
    val, par = DEMC.extract_best(mc)
    assert val == 10.0
    np.testing.assert_array_equal(par, chain[4, :, 1])


def test_julia_shim_structs_match_the_abi_field_for_field():
    """The Julia shim is written without a Julia toolchain in the image, so its `COpts` / `CRunInfo` mirrors are held to
    the ctypes mirror (itself held to the C header by the offset test above): same field names, order and C types."""
    src = open(os.path.join(ROOT, "demc.jl_b200", "julia", "DEMC.jl")).read()
    jl2c = {"Int32": C.c_int32, "Int64": C.c_int64, "UInt64": C.c_uint64, "Float64": C.c_double,
            "Ptr{Int32}": C.POINTER(C.c_int32), "Ptr{Float64}": C.POINTER(C.c_double)}

    def fields(name):
        body = re.search(r"struct %s\n(.*?)\nend" % name, src, flags=re.S).group(1)
        out = []
        for part in re.split(r"[;\n]", body):
            part = part.split("#")[0].strip()
            if part:
                n, t = part.split("::")
                out.append((n.strip(), jl2c[t.strip()]))
        return out

    for jl_name, cls in (("COpts", _lib.Opts), ("CRunInfo", _lib.RunInfo)):
        jl = fields(jl_name)
        py = list(cls._fields_)
        assert [n for n, _ in jl] == [n for n, _ in py], jl_name
        for (n, tj), (_, tp) in zip(jl, py):
            assert C.sizeof(tj) == C.sizeof(tp) and (tj is tp or "LP_" in tp.__name__), (jl_name, n)
    # every ccall names a function the header declares
    called = set(re.findall(r"ccall\(\(:(demcz_[a-z0-9_]+)", src))
    assert called and called <= set(_header_functions())


def test_julia_ccall_signatures_have_the_arity_of_the_c_prototypes():
    hdr = open(os.path.join(ROOT, "include", "demcz.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    proto = {}
    for m in re.finditer(r"\b(demcz_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        args = m.group(2).strip()
        proto[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    src = open(os.path.join(ROOT, "demc.jl_b200", "julia", "DEMC.jl")).read()
    seen = 0
    for m in re.finditer(r"ccall\(\(:(demcz_[a-z0-9_]+),\s*\w+\),\s*\w+,\s*\(([^)]*)\)", src):
        name, types = m.group(1), [t for t in m.group(2).split(",") if t.strip()]
        assert len(types) == proto[name], (name, types, proto[name])
        seen += 1
    assert seen >= 10
