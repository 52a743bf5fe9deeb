"""Philox4x32-10 known-answer tests (Random123 kat_vectors) for the control stream of SURVEY.md 8d, on the oracle
and on the product's device function (host instantiation)."""
import ctypes as C

import numpy as np

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, hostcheck_lib, oracle_lib, philox4x32_10

KATS = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, want in KATS:
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        oracle_lib().oxo_philox4x32_10(c, k, o)
        assert tuple(o) == want
        o2 = (C.c_uint32 * 4)()
        hostcheck_lib().hc_philox(*ctr, *key, o2)
        assert tuple(o2) == want
        assert philox4x32_10(ctr, key) == want      # the pure-Python copy used by tests/test_env_layer.py


def test_control_stream_is_identical_in_fp32_fp64_and_oracle():
    m = ox.Model.from_xml_string(ox.models.HUMANOID)
    nenv = 16
    got = {}
    for prec in ("f32", "f64"):
        hb = HostBatch(m, nenv, prec)
        hb.step(1, True, SEED, 1000, 7)
        got[prec] = hb.get("ctrl")
    assert np.array_equal(got["f32"], got["f64"])          # 2^-23 lattice: exact in both precisions
    assert np.all(np.abs(got["f64"]) < 1) and abs(got["f64"].mean()) < 0.2
    od = OracleData(m)
    for e in range(nenv):
        od.fill_ctrl_philox(1000 + e, 7)
        assert np.array_equal(od.field("ctrl"), got["f64"][e])
    od.fill_ctrl_philox(1000, 8)
    assert not np.array_equal(od.field("ctrl"), got["f64"][0])
