"""CPU tests of the boundary: the C-ABI library builds, loads without a GPU/driver, exports every symbol
include/bmkg_b200.h declares, and the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    import __graft_entry__ as ge

    return ge._load_build_module().build_library()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "bmkg_b200.h")).read()
    return sorted(set(re.findall(r"\b(bmkg_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    names = _declared()
    assert len(names) >= 40
    assert [n for n in names if not hasattr(lib, n)] == []


def test_python_binding_covers_header(libpath):
    from biomedkg_b200 import _cabi

    assert sorted(_cabi.SIGNATURES) == _declared()
    assert _cabi.lib.bmkg_abi_version() == 4
    assert b"workspace" in _cabi.lib.bmkg_error_string(-3)


def test_host_side_size_queries(libpath):
    from biomedkg_b200._cabi import lib

    assert lib.bmkg_infonce_padded_rows(100, 100) == 256 and lib.bmkg_infonce_padded_rows(64, 64) == 128
    assert lib.bmkg_infonce_stacked_rows(1000, 384) == 2304 and lib.bmkg_infonce_padded_rows(1000, 384) == 2304
    assert lib.bmkg_edge_sort_workspace_bytes(1000, 50_000) >= 50_000 * 24
    assert lib.bmkg_csr_filter_workspace_bytes(1000, 50_000) >= 50_000 * 8
    assert lib.bmkg_infonce_workspace_bytes(8000, 256) > 0
    # column phases of the InfoNCE backward (host-side schedule): cfg4 on one GPU = 2032 row blocks x 4 phases, a rank of the
    # 8-GPU row-sharded run = 254 row blocks x 4 phases (7 whole waves of 148 items instead of 1.7 waves of row blocks);
    # N = 28k already fills 2.96 waves with one phase; the phase budget is a process-wide tuning knob
    n4 = 130_000
    assert lib.bmkg_infonce_bwd_workspace_bytes(n4, n4, 256, 0, 2 * n4) == 4 * 2032 * 128 * 257 * 4
    blk = 16_256                                                        # dist.shard_layout(130000, 8): ceil(N/8) rounded up to 128
    assert lib.bmkg_infonce_stacked_rows(n4, blk) == 2 * 8 * blk
    assert lib.bmkg_infonce_bwd_workspace_bytes(n4, blk, 256, 0, 2 * blk) == 4 * 254 * 128 * 257 * 4
    assert lib.bmkg_infonce_bwd_workspace_bytes(28_000, 28_000, 256, 0, 56_000) == 0
    assert lib.bmkg_infonce_bwd_workspace_bytes(n4, n4, 257, 0, 2 * n4) == 0           # unsupported width: nothing to size
    old = lib.bmkg_infonce_set_phase_bytes(1 << 20)
    try:
        assert old == 40 << 20 and lib.bmkg_infonce_set_phase_bytes(0) == 1 << 20    # <= 0 only queries
        assert lib.bmkg_infonce_bwd_workspace_bytes(28_000, 28_000, 256, 0, 56_000) > 0
    finally:
        lib.bmkg_infonce_set_phase_bytes(old)


def test_module_surface_matches_reference(libpath):
    import biomedkg_b200 as b

    m = b.GRACEModule(in_dim=32, hidden_dim=64, out_dim=64, num_hidden_layers=2, scheduler_type="cosine",
                      learning_rate=1e-3, warm_up_ratio=0.2, fuse_method="attention")
    keys = set(m.state_dict())
    for i in range(4):
        assert {f"model.encoder.graph_layers.{i}.lin.weight", f"model.encoder.graph_layers.{i}.bias"} <= keys
    assert {"model.fc1.weight", "model.fc2.bias", "modality_transform.q_proj.weight", "modality_transform.v_proj.bias"} <= keys
    assert m.model.encoder.graph_layers[0].lin.weight.shape == (64, 32)
    assert b.FusionFactory.create_fuser("none", 32) is None and b.FusionFactory.create_fuser(None, 32) is None
    assert type(b.FusionFactory.create_fuser("redaf", 32)).__name__ == "ReDAF"
    opt = m.configure_optimizers.__func__  # noqa: F841 - exists with the reference's name
    d = b.DGIModule(32, 64, 64, 2)
    assert {"model.project.weight", "model.project.bias"} <= set(d.state_dict())
    g = b.GGDModule(32, 64, 64, 2)
    assert {"model.mlp.0.weight", "model.mlp.0.bias"} <= set(g.state_dict())
    # Adam sees self.model only (fuser frozen), as gcl_module.py:81
    m.trainer = type("T", (), {"estimated_stepping_batches": 100})()
    cfg = m.configure_optimizers()
    n_opt = sum(p.numel() for grp in cfg["optimizer"].param_groups for p in grp["params"])
    assert n_opt == sum(p.numel() for p in m.model.parameters())


def test_no_cpu_fallback(libpath):
    import biomedkg_b200 as b

    enc = b.GCNEncoder(32, 64, 64, 2)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        enc(torch.randn(5, 32), torch.zeros(2, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        b.ops.infonce_loss(torch.randn(8, 64), torch.randn(8, 64))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "biomedkg_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_hash_mask_host_mirror():
    from biomedkg_b200.draws import hash_keep_mask

    m = hash_keep_mask(1234, 200_000, 0.2)
    assert abs(float(m.float().mean()) - 0.8) < 5e-3
    assert torch.equal(m, hash_keep_mask(1234, 200_000, 0.2)) and not torch.equal(m, hash_keep_mask(1235, 200_000, 0.2))


@pytest.mark.parametrize("name", ["grace_none", "grace_attention", "dgi_none", "ggd_none_a", "ggd_redaf"])
def test_lightning_checkpoint_layout_loads_reference_state_dict(libpath, golden_dir, name, tmp_path):
    """train_gcl.py:78-83 / node.py:204-209: a checkpoint holds the reference module's state_dict under "state_dict" and
    the kwargs of the outermost __init__ under "hyper_parameters"; load_from_checkpoint must rebuild the module from
    them and load every key strictly (the state_dict here was produced by the reference's own module)."""
    import torch

    import biomedkg_b200 as b

    fx = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)
    cfg = fx["cfg"]
    hp = dict(in_dim=cfg["in_dim"], hidden_dim=cfg["hidden_dim"], out_dim=cfg["out_dim"], num_hidden_layers=cfg["num_hidden_layers"],
              scheduler_type="cosine", learning_rate=2e-4, warm_up_ratio=0.03, fuse_method=cfg["fuse_method"])
    path = str(tmp_path / "epoch=0.ckpt")
    torch.save({"state_dict": {k: v.float() for k, v in fx["state_dict"].items()}, "hyper_parameters": hp,
                "pytorch-lightning_version": "2.2.1", "epoch": 0}, path)
    cls = getattr(b, cfg["cls"])
    mod = cls.load_from_checkpoint(path)
    assert set(mod.state_dict()) == set(fx["state_dict"])
    for k, v in mod.state_dict().items():
        assert torch.equal(v, fx["state_dict"][k].float()), k
    # and our own save -> load round trip keeps the captured hyper-parameters
    mod.save_checkpoint(path)
    again = cls.load_from_checkpoint(path)
    assert {k: again.hparams[k] for k in hp} == hp


def test_header_is_plain_c_and_links(libpath, tmp_path):
    """The boundary is a C ABI: include/bmkg_b200.h must compile as C99 (no C++-isms, no torch types) and a C program must
    link against the shared library and call a host-only entry point (no GPU needed)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "cabi.c"
    src.write_text('#include "bmkg_b200.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%d %lld\\n", bmkg_abi_version(), (long long)bmkg_infonce_padded_rows(100, 100)); return 0; }\n')
    exe = tmp_path / "cabi"
    libdir = os.path.dirname(libpath)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", str(src), "-I", os.path.join(root, "include"), "-L", libdir,
                        "-lbmkg_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split()[1] == "256" and int(out.stdout.split()[0]) >= 1
