"""CPU: the oracle (and the host emulation of the kernel code) against the golden fixtures that
the compiled reference decoder produced (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from common import *  # noqa: F401,F403
import emul


@pytest.mark.parametrize("name", ["qcif_i", "cif_ip", "qcif_wp"])
@pytest.mark.parametrize("impl", ["oracle", "kernel-emulation"])
def test_golden(built, name, impl):
    g = load_golden(name)
    recon = oracle.recon_picture if impl == "oracle" else emul.recon_picture
    outs = run_sequence(recon, g["w_mbs"], g["h_mbs"], abi.REF_COMPAT, g["pics"])
    for i, (p, o) in enumerate(zip(g["pics"], outs)):
        assert o["excursions"] == 0
        for k in "yuv":
            assert np.array_equal(p[k], o[k]), "picture %d plane %s differs from the reference" % (i, k)
        assert planes_md5(o["y"], o["u"], o["v"]) == p["md5"]


def test_golden_1080p_md5(built):
    g = load_golden("hd1080_ip")
    outs = run_sequence(oracle.recon_picture, g["w_mbs"], g["h_mbs"], abi.REF_COMPAT, g["pics"])
    for p, o in zip(g["pics"], outs):
        assert planes_md5(o["y"], o["u"], o["v"]) == p["md5"]
