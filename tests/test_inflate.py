"""The library's own raw-DEFLATE decoder (csrc/inflate_fast.h) behind the native BAM ingest, against zlib:
every block type (stored / fixed / dynamic Huffman), matches of every distance class, BGZF blocks of the
fixture BAMs, truncated and corrupted streams (must be refused or caught by the caller's CRC, never crash)."""
import ctypes
import os
import zlib

import numpy as np
import pytest

from tredparse_b200 import _lib, ingest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lib():
    lib = _lib.load()
    ingest._bind(lib)
    return lib


def raw_deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    return c.compress(data) + c.flush()


def own_inflate(lib, comp, n):
    out = np.empty(max(n, 1), dtype=np.uint8)
    rc = lib.tredsw_inflate_raw(comp, len(comp), out.ctypes.data, n)
    return rc, bytes(out[:n])


def _corpus():
    rng = np.random.default_rng(1234)
    dna = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 70000))
    yield "empty", b""
    yield "one", b"A"
    yield "zeros", bytes(65280)                                   # distance-1 runs (memset path)
    yield "period3", b"CAG" * 20000                               # distance 3: overlapping byte copies
    yield "period7", b"ACGTTGA" * 9000
    yield "period9", b"ACGTTGACC" * 7000                          # distance >= 8: word copies that overlap
    yield "dna", dna
    yield "random", bytes(rng.integers(0, 256, 50000, dtype=np.uint8))   # incompressible: stored / long codes
    yield "text", (b"read_%d\tchr4\t3074877\t60\t150M\t=\t3075100\t373\t" * 900)
    yield "skewed", bytes(np.minimum(rng.geometric(0.02, 65000), 255).astype(np.uint8))   # many code lengths
    with open(os.path.join(GOLDEN, "t001.mini.bam"), "rb") as fp:
        yield "bam_bytes", fp.read(60000)


@pytest.mark.parametrize("name,data", list(_corpus()), ids=[n for n, _ in _corpus()])
def test_every_block_type_equals_zlib(lib, name, data):
    variants = [(lvl, zlib.Z_DEFAULT_STRATEGY) for lvl in (0, 1, 6, 9)]
    variants += [(6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE), (9, zlib.Z_FILTERED)]
    for level, strategy in variants:
        comp = raw_deflate(data, level, strategy)
        assert zlib.decompress(comp, -15) == data
        rc, got = own_inflate(lib, comp, len(data))
        assert rc == 0 and got == data, (name, level, strategy)
        # a wrong expected size is refused
        assert own_inflate(lib, comp, len(data) + 1)[0] != 0
        if len(data) > 1:
            assert own_inflate(lib, comp, len(data) - 1)[0] != 0


def test_multi_block_and_small_windows(lib):
    rng = np.random.default_rng(5)
    data = b"".join(bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), 3000)) + b"CAG" * 500 for _ in range(12))
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = b""
    for i in range(0, len(data), 5000):                           # full flushes: many blocks incl. empty stored ones
        comp += c.compress(data[i:i + 5000]) + c.flush(zlib.Z_FULL_FLUSH)
    comp += c.flush()
    rc, got = own_inflate(lib, comp, len(data))
    assert rc == 0 and got == data
    comp9 = raw_deflate(data, 9, wbits=-9)                         # 512-byte window: short distances only
    assert own_inflate(lib, comp9, len(data)) == (0, data)


def test_truncated_and_corrupted_streams_never_crash(lib):
    rng = np.random.default_rng(9)
    data = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 20000)) + b"CAG" * 3000
    comp = raw_deflate(data)
    for cut in (0, 1, 2, 5, len(comp) // 2, len(comp) - 9, len(comp) - 1):
        rc, got = own_inflate(lib, comp[:cut], len(data))
        assert rc != 0 or got == data                             # refused (a lucky tail may still decode fully)
    refused = wrong = 0
    for k in range(300):
        bad = bytearray(comp)
        pos = int(rng.integers(0, len(bad)))
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        rc, got = own_inflate(lib, bytes(bad), len(data))
        if rc != 0:
            refused += 1
        elif got != data:
            wrong += 1                                            # structurally valid but different: the CRC's job
    assert refused + wrong >= 290
    for k in range(200):                                          # pure noise
        noise = bytes(rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8))
        own_inflate(lib, noise, int(rng.integers(0, 5000)))


@pytest.mark.parametrize("bam", ["t001.mini.bam", "t002.mini.bam"])
def test_bgzf_blocks_of_the_fixtures_use_the_own_decoder(lib, bam):
    from tredparse_b200.meta import TREDsRepo
    repo = TREDsRepo()
    path = os.path.join(GOLDEN, bam)
    with ingest.BamIngest(path) as ing:
        for name in ("HD", "DM1"):
            ing.extract_locus(repo[name], 150, want_names=True)
        own, fallback = ing.inflate_stats()
    assert own > 10 and fallback == 0
    # every BGZF block of the file, decoded by both
    raw = open(path, "rb").read()
    off = nblocks = 0
    while off < len(raw):
        xlen = int.from_bytes(raw[off + 10:off + 12], "little")
        bsize = int.from_bytes(raw[off + 16:off + 18], "little") + 1
        comp = raw[off + 12 + xlen:off + bsize - 8]
        isize = int.from_bytes(raw[off + bsize - 4:off + bsize], "little")
        want = zlib.decompress(comp, -15)
        assert len(want) == isize
        assert own_inflate(lib, comp, isize) == (0, want)
        off += bsize
        nblocks += 1
    assert nblocks > 20
