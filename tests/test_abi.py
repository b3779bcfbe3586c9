"""The C-ABI library loads and exports every symbol include/*.h declares (no compute without a GPU)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared(header):
    txt = (ROOT / "include" / header).read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(esacb200_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("header", ["esac_b200.h", "esac_b200_testhooks.h"])
def test_every_declared_symbol_is_exported(lib, header):
    names = _declared(header)
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/{header} but not exported"


def test_public_header_mirrors_reference_entry_points():
    names = _declared("esac_b200.h")
    assert "esacb200_forward" in names and "esacb200_backward" in names


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert lib.esacb200_create(0, C.byref(h)) != 0  # ESACB200_ERR_NO_DEVICE
    import esac_b200.api as api
    import numpy as np
    from esac_b200.synth import make_scene
    sc = make_scene(E=1, H=8, W=10, M=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        api.forward(sc.coords, sc.assign, np.zeros((4, 4), np.float32), *sc.params)
    # the additive entry points fail the same way: nothing in the product computes on the CPU
    with pytest.raises(RuntimeError, match="no CPU path"):
        api.backward(sc.coords, np.zeros_like(sc.coords), sc.assign, sc.gt_pose, 1.0, 100.0, 100.0, *sc.params)
    with pytest.raises(RuntimeError, match="no CPU path"):
        api.assign_hypotheses(np.ones((1, 3), np.float32), 8, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        api.reproj_loss(sc.coords[:1], sc.gt_pose[None], 525.0, 0, 0, 10.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        api.backward_batch(sc.coords[None], np.zeros_like(sc.coords)[None], sc.assign[None], sc.gt_pose[None], 1.0, 100.0, 100.0,
                           0, 0, *sc.params[2:])


def test_argument_checks_mirror_accessor_errors():
    """at::Tensor::accessor<float,4>() throws RuntimeError on dtype/rank mismatch (esac.cpp:80-84)."""
    import numpy as np
    import esac_b200.api as api
    from esac_b200.synth import make_scene
    sc = make_scene(E=1, H=8, W=10, M=4)
    out = np.zeros((4, 4), np.float32)
    with pytest.raises(RuntimeError, match="expected scalar type Float but found Double"):
        api.forward(sc.coords.astype(np.float64), sc.assign, out, *sc.params)
    with pytest.raises(RuntimeError, match="expected scalar type Long but found Int"):
        api.forward(sc.coords, sc.assign.astype(np.int32), out, *sc.params)
    with pytest.raises(RuntimeError, match="expected 4 dims"):
        api.forward(sc.coords[0], sc.assign, out, *sc.params)


def test_esac_shim_exports_forward_backward():
    import esac
    assert callable(esac.forward) and callable(esac.backward)
