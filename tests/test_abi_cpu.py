"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/kanode.h declares,
validates descriptors, and refuses to compute without a GPU (there is no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import kan_odes_b200 as K
from conftest import lv_chain, source_chain
from kan_odes_b200 import abi

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return abi.load_library()


def test_header_symbols_all_exported(lib):
    hdr = (ROOT / "include" / "kanode.h").read_text()
    declared = set(re.findall(r"\b(kanode_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"kanode_handle"}
    assert declared == set(abi.EXPORTED_SYMBOLS), declared ^ set(abi.EXPORTED_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.kanode_version()


def test_param_count_matches_parameterlength(lib):
    for chain, kw in [(lv_chain(), {}), (lv_chain(40, 5), {}),
                      (source_chain(10), dict(rhs_kind=abi.RHS_SOURCE_LAPLACIAN, n_state=41, lap_coef=-1e-4, dx=0.05))]:
        d = chain.desc(**kw)
        assert lib.kanode_param_count(C.byref(d)) == chain.parameterlength()
    bad = lv_chain().desc(); bad.n_state = 3
    assert lib.kanode_param_count(C.byref(bad)) == 0
    bad = lv_chain().desc(); bad.layers[0].grid_len = 1
    assert lib.kanode_param_count(C.byref(bad)) == 0


def test_desc_struct_layout_matches_header(tmp_path):
    """ctypes mirrors of the descriptor structs against the C compiler's view of include/kanode.h (sizes and field offsets)."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "kanode.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(kanode_layer_desc), sizeof(kanode_desc), sizeof(kanode_stats), offsetof(kanode_layer_desc, kind), '
                   'offsetof(kanode_desc, rhs_kind), offsetof(kanode_desc, lap_coef), offsetof(kanode_desc, dx));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(abi.REPO_ROOT / "include"), "-o", str(exe), str(src)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(abi.LayerDesc), C.sizeof(abi.Desc), C.sizeof(abi.Stats), abi.LayerDesc.kind.offset,
                   abi.Desc.rhs_kind.offset, abi.Desc.lap_coef.offset, abi.Desc.dx.offset]
    assert C.sizeof(abi.LayerDesc) == 44 and C.sizeof(abi.Stats) == 16


def test_create_without_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = lv_chain().desc()
    h = C.c_void_p()
    rc = lib.kanode_create(C.byref(d), 0, None, C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.kanode_last_error(None)
    with pytest.raises(K.KanodeError):
        K.NeuralODE(lv_chain(), (0.0, 3.5), K.Tsit5(), saveat=[0.0])


def test_flat_parameter_layout_roundtrip():
    chain = lv_chain()
    ps, st = K.setup(np.random.default_rng(0), chain)
    p = K.flatten_params(ps)
    assert p.dtype == np.float32 and p.size == 240
    back = K.unflatten_params(chain, p)
    for name in ps:
        for key in ps[name]:
            assert np.array_equal(back[name][key], ps[name][key])
    # column (i-1)*G+g of C belongs to input i (Activation_getter.jl:9-10)
    assert np.array_equal(back["layer_1"]["C"][:, :5], ps["layer_1"]["C"][:, :5])
    assert np.allclose(st["layer_1"]["grid"], [-1, -0.5, 0, 0.5, 1])
    assert chain.layers[0].denominator == 0.5 and chain.layers[0].statelength() == 5


def test_adam_matches_flux_formula():
    """[EXT Flux 0.14.22] Adam: m,v moments with bias correction."""
    rng = np.random.default_rng(0)
    p = rng.normal(size=7); g1, g2 = rng.normal(size=7), rng.normal(size=7)
    opt = K.Adam(5e-4)
    q = p.copy()
    opt.update(q, g1); opt.update(q, g2)
    m = 0.1 * g1; v = 0.001 * g1**2
    ref = p - 5e-4 * (m / (1 - 0.9)) / (np.sqrt(v / (1 - 0.999)) + 1e-8)
    m = 0.9 * m + 0.1 * g2; v = 0.999 * v + 0.001 * g2**2
    ref = ref - 5e-4 * (m / (1 - 0.9**2)) / (np.sqrt(v / (1 - 0.999**2)) + 1e-8)
    assert np.allclose(q, ref, rtol=1e-13)


def test_mat_checkpoint_roundtrip_in_the_reference_layout(tmp_path):
    """`.mat` checkpoint with the reference's variable names and shapes (LV_driver_KANODE.jl:251-272) and its restart read (:149-159)."""
    import numpy as np
    from scipy.io import loadmat
    import kan_odes_b200 as K
    rng = np.random.default_rng(0)
    p_list = [rng.normal(size=240) for _ in range(5)]
    t = np.arange(141) * 0.1
    pred = rng.normal(size=(141, 2))
    f = tmp_path / "LV_kanode_results.mat"
    K.save_checkpoint(f, p_list, [3.0, 2.0, 1.0], [4.0, 3.5, 3.1], t, pred, [2, 10, 5])
    m = loadmat(str(f))
    assert m["p_list"].shape == (5, 240, 1) and m["loss"].shape == (5, 1) and m["kan_pred_u1"].shape == (141, 1)
    assert np.array_equal(np.ravel(m["loss"]), [3.0, 2.0, 1.0, 0.0, 0.0])                # zero padded to len(p_list) like :257-264
    ck = K.load_checkpoint(f)
    assert np.array_equal(ck["p"], p_list[-1]) and ck["size_KAN"] == [2, 10, 5] and np.array_equal(ck["kan_pred_u2"], pred[:, 1])
