"""The C-ABI library loads and exports every symbol include/lshm.h declares (no compute)."""
import ctypes
import os
import re

from lshm_b200 import _lib


def test_header_declares_expected_entry_points():
    protos = _lib.parse_header()
    for name in ("lshm_patchify_scale_i8", "lshm_normalise", "lshm_fft2_reim_shift_clamp", "lshm_down2d", "lshm_up2d",
                 "lshm_wgrad2d", "lshm_down1d", "lshm_up1d", "lshm_wgrad1d", "lshm_linear_fwd", "lshm_cascade_losses",
                 "lshm_khm_fwd", "lshm_khm_bwd", "lshm_khm_assign", "lshm_khm_center_sums", "lshm_similarity",
                 "lshm_augment", "lshm_multiplier_update", "lshm_adam_step", "lshm_last_error", "lshm_fft2_features",
                 "lshm_vec_add", "lshm_down2d_planes", "lshm_wgrad1d_planes", "lshm_cascade_losses_planes",
                 "lshm_tconv_bwd1d_planes", "lshm_tconv_bwd2d_planes", "lshm_tconv_bwd1d", "lshm_tconv_bwd2d"):
        assert name in protos, name
    # every LSHM_API line was understood by the parser
    text = open(_lib.HEADER).read()
    assert len(re.findall(r"^LSHM_API", text, flags=re.M)) == len(protos)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIBRARY), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    cdll = ctypes.CDLL(_lib.LIBRARY)
    for name in _lib.parse_header():
        assert hasattr(cdll, name), f"{name} declared in include/lshm.h but not exported"


def test_binding_loads_and_reports_version():
    L = _lib.lib()
    assert L.version() >= 100
    # argument counts of the binding follow the header
    for name, (_, args) in L.protos.items():
        assert len(getattr(L.cdll, name).argtypes or []) == len(args)


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (tier rule)."""
    pkg = os.path.dirname(_lib.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
