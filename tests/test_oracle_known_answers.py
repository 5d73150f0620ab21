"""CPU: pins the oracle (oracle/flac_oracle.c) against the known-answer vectors in the
reference's own documentation (docs/reference/flac/encode/*.tex, docs/reference/flac/encode.tex)
and against hashlib for MD5."""
import ctypes as C
import hashlib

import numpy as np

import helpers

DOC_SAMPLES = [18, 20, 26, 24, 24, 23, 21, 24, 23, 20]


def test_crc8_doc_vector(built):
    # encode.tex:201-227: header FF F8 C9 18 00 -> C2
    data = bytes.fromhex("FFF8C91800")
    assert helpers.orc().orc_crc8(data, len(data)) == 0xC2
    # running the checksum through itself gives 0
    data += bytes([0xC2])
    assert helpers.orc().orc_crc8(data, len(data)) == 0


def test_crc16_doc_vector(built):
    # encode.tex:316-343
    data = bytes.fromhex("FFF8CC1C00C0EB0000000000000000")
    assert helpers.orc().orc_crc16(data, len(data)) == 0xF093
    data += bytes.fromhex("F093")
    assert helpers.orc().orc_crc16(data, len(data)) == 0


def test_md5_matches_hashlib(built):
    rng = np.random.RandomState(1)
    for n in (0, 1, 55, 56, 57, 63, 64, 65, 1000, 100003):
        data = rng.randint(0, 256, size=n).astype(np.uint8).tobytes()
        out = (C.c_uint8 * 16)()
        helpers.orc().orc_md5(data, n, out)
        assert bytes(out) == hashlib.md5(data).digest()


def test_fixed_order_doc_example(built):
    # encode/fixed.tex:22-52: errors [135, 10, 15, 30, 70] -> order 1
    s = np.array(DOC_SAMPLES, dtype=np.int32)
    err = (C.c_uint64 * 5)()
    order = helpers.orc().orc_best_fixed_order(s.ctypes.data, len(s), err)
    assert list(err) == [135, 10, 15, 30, 70]
    assert order == 1


def test_residual_partition_doc_example(built):
    # encode/residual.tex:29-61 works this example with Rice = floor(log2(sum / n)) and gets 38 vs 42.
    # The C code (flac.c:1477-1501) instead grows k while (n << k) < sum, which gives k = 2 where
    # the text has 1 (9 << 1 = 18 < 20): sizes 37 vs 19 + 22 = 41.  The oracle follows the C code
    # (tests/test_oracle_vs_ref.py pins that against the compiled reference); the conclusion of
    # the worked example -- partition order 0 wins -- is the same.
    r = np.array([2, 6, -2, 0, -1, -2, 3, -1, -3], dtype=np.int32)
    rice = (C.c_uint8 * 2)()
    plen = (C.c_uint * 2)()
    t0 = helpers.orc().orc_residual_partitions(r.ctypes.data, 9, 10, 1, 0, 14, rice, plen)
    assert (t0, rice[0], plen[0]) == (37, 2, 9)
    t1 = helpers.orc().orc_residual_partitions(r.ctypes.data, 9, 10, 1, 1, 14, rice, plen)
    assert t1 == 41 and list(rice) == [2, 1] and list(plen) == [4, 5]
    assert t0 < t1


def test_tukey_window_and_autocorrelation_doc_example(built):
    # encode/lpc.tex:30-48 (window, hand-rounded to 2 digits) and :58-141,173 (autocorrelation)
    w = np.zeros(10, dtype=np.float64)
    helpers.orc().orc_tukey_window(10, w.ctypes.data)
    assert np.allclose(w, [0.00, 0.41, 0.97, 1, 1, 1, 1, 0.97, 0.41, 0.00], atol=0.006)
    windowed = np.array([0.0, 8.2, 25.2, 24.0, 24.0, 23.0, 21.0, 23.3, 9.4, 0.0])
    autoc = np.zeros(4)
    helpers.orc().orc_autocorrelate(3, windowed.ctypes.data, 10, autoc.ctypes.data)
    assert np.allclose(autoc, [3455.53, 3130.76, 2590.90, 2055.80], atol=0.01)


def test_levinson_doc_example(built):
    # encode/lpc.tex:173-193
    autoc = np.array([3455.53, 3130.76, 2590.90, 2055.80])
    lp = np.zeros(32 * 32)
    err = np.zeros(32)
    helpers.orc().orc_lp_coefficients(3, autoc.ctypes.data, lp.ctypes.data, err.ctypes.data)
    assert np.allclose(lp[0:1], [0.906], atol=1e-3)
    assert np.allclose(lp[32:34], [1.266, -0.397], atol=2e-3)
    assert np.allclose(lp[64:67], [1.28, -0.443, 0.036], atol=2e-3)
    assert np.allclose(err[:3], [619.107, 521.530, 520.854], rtol=2e-3)


def test_order_estimate_and_quantisation_doc_example(built):
    # encode/lpc.tex:205-228: best order 1; :240-268: [1311, -454, 37], shift 10
    err = np.array([619.107, 521.530, 520.854])
    assert helpers.orc().orc_estimate_best_lpc_order(16, 12, 3, 10, err.ctypes.data) == 1
    c = np.array([1.280, -0.443, 0.036])
    q = (C.c_int * 3)()
    shift = C.c_int()
    helpers.orc().orc_quantize_coefficients(c.ctypes.data, 3, 12, q, C.byref(shift))
    assert list(q) == [1311, -454, 37] and shift.value == 10


def test_doc_sample_frame_bytes(built):
    # the compiled reference encodes the documentation's 10 samples
    # (-c 1 -B 10 -l 3 -R 6) to this frame, whose subframe bits equal figures/fixed-enc-example.bpx
    pcm = helpers.pack_pcm(np.array(DOC_SAMPLES, dtype=np.int32), 16)
    flac = helpers.oracle_encode(pcm, 44100, 1, 16, helpers.options(block_size=10, max_lpc_order=3,
                                                                   max_residual_partition_order=6))
    ff = helpers.first_frame_offset(flac)
    assert flac[ff:].hex() == "fff8690800092212001200904f2f6aa0614a"
    assert ff == 4183 and len(flac) == 4201


def test_rice_parameter_uint32_wrap(built):
    # SURVEY.md H3, confirmed there on the reference's own function: constant-magnitude residuals,
    # max k = 30.  plength 4096: |r| = 2^16 -> k 16, 2^19 -> 19, >= 2^20 -> 30 (wrapped shift)
    def k_for(plength, mag):
        r = np.full(plength, mag, dtype=np.int32)
        rice = (C.c_uint8 * 1)()
        helpers.orc().orc_residual_partitions(r.ctypes.data, plength, plength, 0, 0, 30, rice, None)
        return rice[0]
    assert k_for(4096, 1 << 16) == 16
    assert k_for(4096, 1 << 19) == 19
    assert k_for(4096, 1 << 20) == 30
    assert k_for(4084, 1 << 20) == 20
    assert k_for(4084, 1 << 21) == 30
    assert k_for(1024, 1 << 22) == 30
    assert k_for(64, 1 << 24) == 24


def test_wasted_bits(built):
    s = np.array([0, 8, -16, 24, 0], dtype=np.int32)
    assert helpers.orc().orc_wasted_bits(s.ctypes.data, 5) == 3
    z = np.zeros(7, dtype=np.int32)
    assert helpers.orc().orc_wasted_bits(z.ctypes.data, 7) == 0
    o = np.array([4, 6, 1], dtype=np.int32)
    assert helpers.orc().orc_wasted_bits(o.ctypes.data, 3) == 0
