"""CPU: pin the oracle (C restatement) against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py) and, when it was built, against the reference itself."""
import hashlib

import numpy as np
import pytest

import oracle
import synth


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_keystream_first64(golden):
    for k, hexbytes in golden["keystream_first64"].items():
        assert oracle.keystream(int(k, 16), 64).tobytes().hex() == hexbytes, k


def test_survey_known_answers():
    # SURVEY.md section 8(c), captured there from the unmodified reference
    assert oracle.keystream(synth.PS3_KEY, 16).tobytes().hex(" ") == "47 c6 7c 86 1d 86 ae 50 8a de e9 7f 58 b9 9d 0b"
    assert oracle.keystream(synth.PS4_KEY, 16).tobytes().hex(" ") == "7a cc ad 6f af 91 a7 e3 72 00 8f 07 19 ba 34 03"
    assert oracle.keystream(1, 16).tobytes().hex(" ") == "58 0e 26 d5 7d 37 27 01 bc b2 67 aa 73 1d 4c b8"
    assert oracle.keystream(1, 10000)[9999] == 0xEE  # Park-Miller check value x_10000 = 1043618065
    for k in (0, 0x7FFFFFFF):
        assert not oracle.keystream(k, 256).any()  # identity stream
    for k in (0x80000000, 0xFFFFFFFF, 0x7FFFFFFE):
        assert oracle.keystream(k, 8).tobytes().hex(" ") == "a7 f1 d9 2a 82 c8 d8 fe"
    assert oracle.lib().oracle_keystream_byte(synth.i32(synth.PS4_KEY), 1 << 30) == 0x33


def test_park_miller_check_value():
    # the 10000th state from seed 1 is the published minimal-standard check value
    k = 1
    for _ in range(10000):
        k = oracle.lib().oracle_cycle_key(k)
    assert k == 1043618065


def test_closed_form_matches_deep_bytes(golden):
    for k, table in golden["deep_bytes"].items():
        key = synth.i32(int(k, 16))
        for pos, val in table.items():
            if pos.startswith("window"):
                got = oracle.cycle_at(np.zeros(32, np.uint8), key, 1 << 29)
                assert got.tobytes().hex() == val
            else:
                assert oracle.lib().oracle_keystream_byte(key, int(pos)) == val, (k, pos)


def test_roundtrips(golden):
    for case in golden["roundtrip"]:
        plain = synth.payload(0, case["size"])
        assert sha(plain) == case["plain_sha256"]
        enc = oracle.cycle(plain, int(case["key"], 16))
        assert sha(enc) == case["cycled_sha256"]
        assert enc[:16].tobytes().hex() == case["cycled_head"]
        assert (oracle.cycle(enc, int(case["key"], 16)) == plain).all()  # involution


def test_unaligned_windows(golden):
    big = synth.payload(0, 1 << 16)
    for case in golden["unaligned"]:
        buf = big.copy()
        s, n = case["start"], case["size"]
        buf[s:s + n] = oracle.cycle(buf[s:s + n], int(case["key"], 16))
        assert sha(buf) == case["buffer_sha256"]


def test_batch_fixture(golden):
    b = golden["batch"]
    descs = np.zeros(len(b["len"]), dtype=oracle.DESC_DTYPE)
    descs["src_off"], descs["dst_off"], descs["len"], descs["key"] = b["src_off"], b["dst_off"], b["len"], b["key"]
    src = synth.payload(0, b["src_bytes"])
    dst = np.full(b["dst_bytes"], b["dst_fill"], dtype=np.uint8)
    oracle.cycle_batch(descs, src, dst)
    assert sha(dst) == b["dst_sha256"]


def test_jump_consistency():
    rng = np.random.default_rng(5)
    for key in synth.EDGE_KEYS + [int(x) for x in rng.integers(0, 1 << 32, size=8)]:
        full = oracle.keystream(key, 5000)
        for pos in (0, 1, 15, 16, 17, 1234, 4999):
            kj = oracle.key_jump(key, pos)
            assert (oracle.keystream(kj, 5000 - pos) == full[pos:]).all(), (hex(key), pos)
    # exponent reduction: the stream has period m-1
    assert oracle.key_jump(12345, 0x7FFFFFFE) == 12345
    assert oracle.pow_a(0x7FFFFFFE) == 1


@pytest.mark.skipif(not oracle.have_ref(), reason="unmodified reference not built on this box")
def test_oracle_equals_unmodified_reference():
    rng = np.random.default_rng(11)
    keys = synth.EDGE_KEYS + [int(x) for x in rng.integers(0, 1 << 32, size=24)]
    for key in keys:
        data = synth.payload(key & 0xFFFF, 200_000)
        assert (oracle.cycle(data, key) == oracle.cycle(data, key, use_ref=True)).all(), hex(key)


@pytest.mark.skipif(not oracle.have_ref(), reason="unmodified reference not built on this box")
def test_ref_parts_fanout_equals_serial():
    buf = synth.payload(0, 300_000)
    parts = np.zeros(7, dtype=oracle.PART_DTYPE)
    parts["off"] = [0, 10, 5000, 70000, 70001, 150000, 299999]
    parts["len"] = [10, 4990, 65000, 1, 79999, 149999, 1]
    parts["key"] = synth.entry_keys(7)
    want = buf.copy()
    for p in parts:
        o, l = int(p["off"]), int(p["len"])
        want[o:o + l] = oracle.cycle(want[o:o + l], int(p["key"]), use_ref=True)
    got = buf.copy()
    oracle.ref_cycle_parts(got, parts, 4)
    assert (got == want).all()
