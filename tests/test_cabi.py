"""CPU tier: the C-ABI library builds for sm_100a, loads, and exports every symbol include/aesr_b200.h declares.
No compute calls here (no GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "aesr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aesr_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from superresolution_aniso_mri_b200 import _lib, build
    path = build.build_library()
    assert os.path.exists(path)
    lib = _lib.load()
    declared = header_symbols()
    assert declared == _lib.exported_symbols(), "ctypes signature table out of sync with include/aesr_b200.h"
    for name in declared:
        assert getattr(lib, name) is not None


def test_header_cites_reference_lines():
    src = open(os.path.join(ROOT, "include", "aesr_b200.h")).read()
    for cite in ("networks/acai_vanilla.py:51", "networks/acai_vanilla.py:98", "generate_hr_volumes.py:88",
                 "lpips/pretrained_networks.py:107-116"):
        assert cite in src


def test_sass_has_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, TMA -> UTMALDG, tcgen05.ld -> LDTM (B200_PROFILING.md, 'What proves a Blackwell kernel')."""
    import shutil
    import subprocess
    from superresolution_aniso_mri_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.build_library()], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HGMMA" not in sass


def test_product_path_refuses_cpu_tensors():
    """No CPU fallback: operators raise on non-CUDA tensors instead of computing something else."""
    import torch
    from superresolution_aniso_mri_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ops.e0(torch.zeros(1, 1, 8, 8), torch.zeros(32), torch.zeros(32))


def test_model_state_dict_layout_and_init_stream():
    """VanillaACAI holders reproduce the reference's state_dict keys/shapes and RNG stream (oracle-pinned init)."""
    import torch
    from oracle import aesr_oracle as O
    from superresolution_aniso_mri_b200.networks.acai_vanilla import VanillaACAI
    for lw in (32, 16):
        args = O.default_args(128, lw)
        torch.manual_seed(892372)
        m = VanillaACAI(dict(args))
        want = O.init_state(args, seed=892372)
        sd = m.state_dict()
        assert list(sd.keys()) == list(want.keys())
        assert all(torch.equal(sd[k], want[k]) for k in sd)
    with pytest.raises(RuntimeError):
        m.enc[1](torch.zeros(1))         # holders never compute


def test_top_level_dropin_import_paths():
    """settings.yaml names modules by path (kwatsch/get_trainer.py:61-78): those dotted paths must resolve here."""
    import importlib
    mod = importlib.import_module("networks.acai_vanilla")
    assert hasattr(mod, "VanillaACAI")
    ghv = importlib.import_module("generate_hr_volumes")
    assert hasattr(ghv, "create_super_volume") and hasattr(ghv, "latent_space_interp")


def test_synthesis_plan_tables_match_per_volume_pair_plan():
    """The cached device-resident index / weight tables of a batched synthesis problem (synthesis._synthesis_plan) equal
    the per-volume plan (pair_plan: out[i*(A+1)+1+k] = dec(w_hi[k] z[i+1] + w_lo[k] z[i]), generate_hr_volumes.py:46-66),
    including the degenerate shapes (one slice, no alphas)."""
    import numpy as np
    import torch
    from superresolution_aniso_mri_b200 import synthesis as S
    for V, Z, ar in ((2, 1, [0.5]), (1, 2, []), (3, 4, [0.25, 0.5, 0.75]), (1, 2, [0.5]), (5, 10, list(np.linspace(0, 1, 8)[1:-1]))):
        p = S._synthesis_plan(V, Z, ar, torch.device("cpu"))
        A = len(ar)
        Zo = (Z - 1) * (A + 1) + 1
        assert p["idx"].tolist() == list(range(V * Z))
        assert p["oi_kept"].tolist() == [v * Zo + i * (A + 1) for v in range(V) for i in range(Z)]
        assert p["pa"].numel() == V * (Z - 1) and p["oi"].numel() == V * (Z - 1) * A
        if Z > 1 and A > 0:
            w_hi, w_lo = S.interp_weights(ar)
            assert np.array_equal(p["wa"].numpy(), w_hi) and np.array_equal(p["wb"].numpy(), w_lo)
            for v in range(V):
                ia, ib, _, _, oi = S.pair_plan(Z, A, w_hi, w_lo, "cpu", slice_offset=v * Z, out_offset=v * Zo)
                q0 = v * (Z - 1)
                assert np.array_equal(p["pa"][q0:q0 + Z - 1].numpy(), ia[::A])
                assert np.array_equal(p["pb"][q0:q0 + Z - 1].numpy(), ib[::A])
                assert np.array_equal(p["oi"][q0 * A:(q0 + Z - 1) * A].numpy(), oi)
    assert S._synthesis_plan(3, 4, [0.25, 0.5, 0.75], torch.device("cpu")) is S._synthesis_plan(3, 4, [0.25, 0.5, 0.75],
                                                                                                 torch.device("cpu"))
