"""CPU: the C-ABI library loads and exports exactly what include/tokred.h declares; the ctypes binding mirrors the
header; ops trace with FakeTensors; the product path fails loudly without CUDA (no compute calls here)."""
import os
import re
from argparse import Namespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tokred.h")


def declared():
    """{name: number of parameters} parsed from the header."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"TOKRED_API\s+[\w\s\*]+?\b(tokred_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("void", "") else len([p for p in params.split(",") if p.strip()])
    return out


def test_build_entry_compiles_library():
    import __graft_entry__ as ge
    ge.build()


def test_library_exports_every_declared_symbol():
    from tokenreduction_b200 import _lib
    lib = _lib.load()
    decl = declared()
    assert len(decl) >= 19
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/tokred.h but not exported"
    assert set(_lib.EXPORTS) == set(decl), set(_lib.EXPORTS) ^ set(decl)
    assert lib.tokred_abi_version() == _lib.ABI_VERSION
    assert isinstance(lib.tokred_last_error(), bytes)
    assert _lib.launch_count() >= 0


def test_binding_arity_matches_header():
    from tokenreduction_b200 import _lib
    decl = declared()
    for name, argtypes in _lib.SIGNATURES.items():
        assert len(argtypes) == decl[name], f"{name}: binding has {len(argtypes)} params, header {decl[name]}"


def test_host_side_argument_errors_need_no_gpu():
    """argument validation happens on the host before any launch: negative return + message, no CUDA call."""
    from tokenreduction_b200 import _lib
    lib = _lib.load()
    assert lib.tokred_tome_effective_r(197, 59, 1) == 59
    assert lib.tokred_tome_effective_r(197, 150, 1) == 98
    assert lib.tokred_tome_effective_r(3, 5, 1) == 1
    with pytest.raises(_lib.TokredError, match="null tensor"):
        _lib.call("tokred_topk_gather", None, 0, None, 0, 1, 196, None, 0, 0, 2, 197, 64, 10, None, None, None)
    with pytest.raises(_lib.TokredError, match="argument"):
        _lib.call("tokred_dpcknn_cluster", 16, 0, 16, 2, 196, 64, 300, 5, 0, 16, 16, None)      # K > P
    with pytest.raises(_lib.TokredError, match="x_batch_stride"):
        _lib.call("tokred_dpcknn_cluster", 16, 100, 16, 2, 196, 64, 49, 5, 0, 16, 16, None)     # images overlap
    # empty batch is a no-op even with null tensors
    _lib.call("tokred_topk_gather", None, 0, None, 0, 1, 196, None, 0, 0, 0, 197, 64, 10, None, None, None)


def test_ops_have_no_cpu_path():
    import tokenreduction_b200.ops as T
    with pytest.raises((NotImplementedError, RuntimeError)):
        T.topk_gather(torch.zeros(2, 197, 64), torch.zeros(2, 196), 10)
    with pytest.raises((NotImplementedError, RuntimeError)):
        T.tome_match(torch.zeros(2, 197, 64), 59, True, False)


def test_fake_tensor_shapes():
    from torch._subclasses.fake_tensor import FakeTensorMode
    import tokenreduction_b200.ops as T
    with FakeTensorMode():
        x = torch.empty(4, 197, 384, device="cuda")
        out, idx = T.topk_gather(x, torch.empty(4, 196, device="cuda"), 137)
        assert out.shape == (4, 138, 384) and idx.shape == (4, 137) and idx.dtype == torch.int64
        out, idx, compl = T.evit_select_fuse(x, torch.empty(4, 196, device="cuda"), 98)
        assert out.shape == (4, 100, 384) and idx.shape == (4, 99) and compl.shape == (4, 98)
        unm, src, dst = T.tome_match(torch.empty(4, 197, 64, device="cuda"), 150, True, True)
        assert unm.shape == (4, 1) and src.shape == (4, 98)          # r capped at (N-1)//2 (quirk B.4)
        o, s, m = T.tome_merge(x, None, unm, src, dst, True, True)
        assert o.shape == (4, 99, 384) and s.shape == (4, 99, 1) and m.shape == (4, 196)
        xp = torch.empty(4, 196, 384, device="cuda")
        ic, idn = T.dpcknn_cluster(xp, torch.empty(4, 196, device="cuda"), 49, 5)
        assert ic.shape == (4, 196) and idn.shape == (4, 49)
        c, ci, a = T.kmedoids_fit(xp, torch.empty(4, 196, 1, device="cuda"), 49, 3)
        assert c.shape == (4, 49, 384) and ci.shape == (4, 49) and a.shape == (4, 196)
        o, w = T.sinkhorn_merge(xp, torch.empty(176, 384, device="cuda"), 1.0, 3, True)
        assert o.shape == (4, 176, 384) and o.dtype == torch.bfloat16 and w.shape == (4, 176, 196)
        ids, mk, mc = T.ats_sample(torch.empty(4, 6, 197, 64, device="cuda"), torch.empty(4, 6, 197, 197, device="cuda"),
                                   torch.empty(4, 197, dtype=torch.bool, device="cuda"), torch.empty(176, device="cuda"))
        assert ids.shape == (4, 177) and mk.dtype == torch.bool and mc.shape == (1,)


def test_factory_surface():
    """the 42 reference entrypoint names exist; reduced models keep the helper methods and parameter names."""
    import contextlib
    import io
    from tokenreduction_b200 import create_model, list_models
    names = list_models()
    assert len(names) == 42
    for fam in ("topk", "evit", "tome", "dyvit", "dpcknn", "kmedoids", "sinkhorn", "patchmerger", "ats", "sit", "heuristic"):
        for size in ("tiny", "small", "base"):
            assert f"{fam}_{size}_patch16_224" in names
    args = Namespace(keep_rate=[0.25], reduction_loc=[3, 6, 9], k_neighbors=5, equal_weight=False)
    with contextlib.redirect_stdout(io.StringIO()):
        m = create_model("dpcknn_tiny_patch16_224", pretrained=False, num_classes=7, drop_rate=0.0, drop_path_rate=0.0,
                         drop_block_rate=None, img_size=224, args=args)
    assert m.get_new_module_names() == ["cluster_layers"] and m.get_reduction_count() == [3, 6, 9]
    assert m.cluster_count == [49, 12, 3]
    assert "cluster_layers.0.score.weight" in m.state_dict() and m.head.out_features == 7
    with pytest.raises(RuntimeError):
        create_model("no_such_model")
    with pytest.raises(NotImplementedError):
        create_model("heuristic_small_patch16_224", args=args)


def test_keep_rate_schedules():
    import contextlib
    import io
    from tokenreduction_b200 import create_model
    mk = lambda name, kr: create_model(name, args=Namespace(keep_rate=[kr], reduction_loc=[3, 6, 9], k_neighbors=5,
                                                            cluster_iters=3, sinkhorn_eps=1.0, equal_weight=False,
                                                            dyvit_distill=False, distillation_type="none"))
    with contextlib.redirect_stdout(io.StringIO()):
        tome = mk("tome_tiny_patch16_224", 0.7)
        ats = mk("ats_tiny_patch16_224", 0.9)
        topk = mk("topk_tiny_patch16_224", 0.7)
    assert [b.r for b in tome.blocks] == [0, 0, 0, 59, 0, 0, 41, 0, 0, 29, 0, 0]
    assert ats.sample_count == [0, 0, 0, 177, 0, 0, 159, 0, 0, 143, 0, 0]
    assert [round(b.attn.keep_rate, 3) for b in topk.blocks][3::3] == [0.7, 0.49, 0.343]


def test_bench_byte_formulas_follow_the_header():
    """bench.py addresses ABI arguments by the parameter names of include/tokred.h: every entry point must parse to
    as many names as the ctypes signature has arguments, and two known per-launch figures must come out
    (ToMe merge stage 1 at DeiT-S fp32 = 517,160 B/image)."""
    import bench
    from tokenreduction_b200 import _lib
    names = bench.abi_arg_names()
    for fn, sig in _lib.SIGNATURES.items():
        assert len(names[fn]) == len(sig), fn
    args = (1, 0, None, 1, 1, 1, 256, 197, 384, 59, 1, 1, 1, 1, None)       # x, fp32, size=NULL (first stage), ..., rci, divide, stream
    assert bench.algorithmic_bytes("tokred_tome_merge", args) == 256 * 517160
    args = (1, 0, 197 * 768, 1, 128, 196, 768, 176, 1.0, 0.0, 3, 1, 1, 1, 1, None, 0, None)   # sinkhorn, fp32 x, bf16 out
    assert bench.algorithmic_bytes("tokred_sinkhorn_merge", args) == 128 * (196 * 768 * 4 + 176 * 768 * 2 + 176 * 196 * 4) + 176 * 768 * 4


def test_rows_view_helper_host_logic():
    """ops._rows: which token tensors are handed to the kernels in place (pointer + x_batch_stride) and which are
    copied.  Pure host logic, no launch."""
    import tokenreduction_b200.ops as T
    full = torch.zeros(4, 197, 64)
    x, s = T._rows(full)
    assert x is full and s == 0                                    # dense
    v = full[:, 1:]
    x, s = T._rows(v)
    assert x.data_ptr() == v.data_ptr() and s == 197 * 64          # class token dropped: read in place
    x, s = T._rows(full[:1, 1:])
    assert s == 0                                                  # a single image needs no stride
    x, s = T._rows(full[:, :, :32])
    assert x.is_contiguous() and s == 0                            # rows not dense: copied
    x, s = T._rows(full.transpose(1, 2)[:, :, :64].transpose(1, 2)[:, ::2])
    assert x.is_contiguous() and s == 0                            # row stride != C: copied
    x, s = T._rows(full[::2, 1:])
    assert x.data_ptr() == full[::2, 1:].data_ptr() and s == 2 * 197 * 64


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """include/tokred.h must be consumable by a C compiler (the boundary is a C ABI, not C++), and a C program linked
    against the library must be able to call it: version query, host-side validation, error string.  No GPU needed."""
    import shutil
    import subprocess
    from tokenreduction_b200 import _lib
    from tokenreduction_b200.build import LIB_PATH
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    _lib.load()
    src = tmp_path / "consumer.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "tokred.h"\n'
        "int main(void) {\n"
        "  if (tokred_abi_version() != TOKRED_ABI_VERSION) return 2;\n"
        "  if (tokred_tome_effective_r(197, 59, 1) != 59) return 3;\n"
        "  /* K > P: rejected on the host, nothing launched */\n"
        "  int rc = tokred_dpcknn_cluster((const float*)16, 0, (const float*)16, 2, 196, 64, 300, 5, 0, (int64_t*)16, (int64_t*)16, 0);\n"
        "  if (rc >= 0) return 4;\n"
        "  if (!strstr(tokred_last_error(), \"cluster_num\")) return 5;\n"
        '  printf("abi %d ok\\n", tokred_abi_version());\n'
        "  return 0;\n}\n")
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(LIB_PATH)
    inc = os.path.join(os.path.dirname(libdir), "include")
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe), "-L", libdir,
           "-l:" + os.path.basename(LIB_PATH), "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "abi" in r.stdout, (r.returncode, r.stdout, r.stderr)
