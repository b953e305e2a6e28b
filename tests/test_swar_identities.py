"""Arithmetic identities the 16-byte block scans rely on (flowgger_b200/csrc/fg_rfc5424.cuh: swar_eq, swar_top_nibble,
first_hit16, swar_json_stop), restated in numpy and checked exhaustively / on random blocks against a naive byte loop.
The device code itself is covered by the GPU parity tests; this pins the claims made in its comments:
  * swar_eq is EXACT per byte (no borrow artefacts, unlike the cheap haszero form),
  * one multiply gathers the four 0x80 flags of a word into the top nibble and no cross term carries into it,
  * first_hit16 returns the first flagged byte at or after the cursor, else the start of the next block.
"""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def swar_eq(w, pat):
    x = (w ^ pat) & M32
    return (~((((x & np.uint64(0x7F7F7F7F)) + np.uint64(0x7F7F7F7F)) | x | np.uint64(0x7F7F7F7F)))) & M32


def top_nibble(z):
    return ((z * np.uint64(0x00204081)) & M32) >> np.uint64(28)


def first_hit16(z, sh, i):
    m = top_nibble(z[0]) | (top_nibble(z[1]) << np.uint64(4)) | (top_nibble(z[2]) << np.uint64(8)) | (top_nibble(z[3]) << np.uint64(12))
    m = int(m) & (0xFFFFFFFF << sh) & 0xFFFFFFFF
    hit = m != 0
    mm = m | 0x10000
    ffs = (mm & -mm).bit_length()  # 1-based index of the lowest set bit
    return i + (ffs - 1) - sh, hit


def test_swar_eq_exact_for_every_byte_pair():
    vals = np.arange(256, dtype=np.uint64)
    for lane in range(4):
        for pat_byte in (0x09, 0x20, 0x22, 0x3A, 0x5C, 0x00, 0xFF, 0x80, 0x7F):
            pat = np.uint64(pat_byte * 0x01010101)
            # neighbours chosen to provoke borrows/carries in the cheap form: pat, pat+1, 0x00, 0xFF
            for nb in (pat_byte, (pat_byte + 1) & 0xFF, 0x00, 0xFF):
                w = np.uint64(nb * 0x01010101) & ~(np.uint64(0xFF) << np.uint64(8 * lane)) | (vals << np.uint64(8 * lane))
                z = swar_eq(w, pat)
                for j in range(4):
                    byte = (w >> np.uint64(8 * j)) & np.uint64(0xFF)
                    flag = (z >> np.uint64(8 * j)) & np.uint64(0xFF)
                    assert np.array_equal(flag, np.where(byte == pat_byte, 0x80, 0).astype(np.uint64))


def test_top_nibble_gathers_all_sixteen_flag_sets():
    for bits in range(16):
        z = np.uint64(sum(0x80 << (8 * j) for j in range(4) if bits >> j & 1))
        assert int(top_nibble(z)) == bits


def test_json_stop_mask():
    # '"', '\\' and every control byte < 0x20 stop the scan; nothing else does (serde_json read.rs ESCAPE table)
    vals = np.arange(256, dtype=np.uint64)
    w = vals * np.uint64(0x01010101)
    z = swar_eq(w, np.uint64(0x22222222)) | swar_eq(w, np.uint64(0x5C5C5C5C)) | swar_eq(w & np.uint64(0xE0E0E0E0), np.uint64(0))
    want = np.where((vals == 0x22) | (vals == 0x5C) | (vals < 0x20), 0x80808080, 0).astype(np.uint64)
    assert np.array_equal(z, want)


def test_first_hit16_matches_a_byte_loop():
    rng = np.random.default_rng(7)
    stops = (0x09, 0x3A)
    for _ in range(20000):
        # dense in stop bytes and their +1 neighbours so that several hits and borrow patterns share a word
        block = rng.choice(np.array([0x09, 0x0A, 0x3A, 0x3B, 0x41, 0x00, 0xFF], dtype=np.uint8), size=16,
                           p=[0.08, 0.08, 0.08, 0.08, 0.6, 0.04, 0.04])
        sh = int(rng.integers(0, 16))
        i = int(rng.integers(0, 1000))
        words = block.view("<u4").astype(np.uint64)
        z = [swar_eq(w, np.uint64(0x09090909)) | swar_eq(w, np.uint64(0x3A3A3A3A)) for w in words]
        got, hit = first_hit16(z, sh, i)
        want = next((k for k in range(sh, 16) if block[k] in stops), None)
        if want is None:
            assert not hit and got == i + 16 - sh
        else:
            assert hit and got == i + want - sh
