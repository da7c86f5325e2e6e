"""gzip on the device (gtgpu_gunzip and the *_gz ingest entry points) against Python's zlib: every DEFLATE block type,
BGZF files, corrupt members, and text that goes straight from the inflater into the device parser."""
import gzip
import struct
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from gtars_b200 import ffi
    c = ffi.Context(0)
    yield c
    c.close()


def bgzf(data: bytes, block=60_000, level=6) -> bytes:
    """A BGZF file as bgzip writes it: independent gzip members with the 'BC' extra field (BSIZE) + the empty EOF block."""
    out = []
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + [b""]
    for ch in chunks:
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = co.compress(ch) + co.flush()
        bsize = 12 + 6 + len(body) + 8
        out.append(b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
                   + body + struct.pack("<II", zlib.crc32(ch), len(ch)))
    return b"".join(out)


def _texts():
    rng = np.random.default_rng(2026)
    bed = "".join(f"chr{rng.integers(1, 23)}\t{a}\t{a + int(rng.integers(1, 900))}\tname{a % 97}\t{a % 1000}\t+\n"
                  for a in rng.integers(0, 200_000_000, 40_000)).encode()
    return {
        "empty": b"",
        "one_byte": b"x",
        "tiny": b"chr1\t10\t20\n",                                      # fixed-Huffman block
        "bed": bed,                                                      # dynamic blocks, matches near and far
        "runs": b"a" * 70_000 + b"ab" * 40_000 + b"abc" * 30_000,        # matches that overlap themselves (distance 1, 2, 3)
        "random": rng.integers(0, 256, 300_000, dtype=np.uint8).tobytes(),  # incompressible: stored blocks
        "long_codes": bytes(rng.choice(256, 400_000, p=np.r_[np.full(8, 0.1), np.full(248, 0.2 / 248)]).astype(np.uint8)),  # skewed: > 10-bit codes
    }


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_gunzip_single_members_match_zlib(ctx, level):
    from gtars_b200 import ffi
    texts = _texts()
    # every text as its own gzip member, all members back to back in one call (= a batch of .gz files)
    blobs = [gzip.compress(t, compresslevel=level, mtime=0) for t in texts.values()]
    # header variants: FNAME + FCOMMENT + FHCRC + FEXTRA that is not BGZF
    body = blobs[3][10:]
    fancy = (b"\x1f\x8b\x08" + bytes([4 | 8 | 16]) + b"\0" * 6 + struct.pack("<H", 5) + b"XYabc" + b"file.bed\0" + b"a comment\0" + body)
    blobs.append(fancy)
    want = list(texts.values()) + [list(texts.values())[3]]
    gz = b"".join(blobs)
    mo = np.concatenate([[0], np.cumsum([len(b) for b in blobs])]).astype(np.uint64)
    text, off = ffi.gunzip(ctx, gz, mo)
    assert list(off) == list(np.concatenate([[0], np.cumsum([len(t) for t in want])]))
    for k, t in enumerate(want):
        assert text[int(off[k]):int(off[k + 1])] == t, (level, k)
    # laid out at odd offsets inside a larger buffer (members need no alignment)
    pad = b"\x07" * 3
    gz2 = pad + blobs[3] + pad + blobs[2]
    mo2 = np.array([3, 3 + len(blobs[3])], dtype=np.uint64)
    text2, _ = ffi.gunzip(ctx, gz2, mo2)
    assert text2 == want[3]


def test_bgzf_members_are_split_and_inflated_in_parallel(ctx):
    from gtars_b200 import ffi
    data = _texts()["bed"] * 3
    gz = bgzf(data)
    mo = ffi.gzip_members(gz)
    assert len(mo) - 1 == -(-len(data) // 60_000) + 1                    # data blocks + the EOF block
    assert gzip.decompress(gz) == data                                    # the fixture is a valid gzip file
    text, off = ffi.gunzip(ctx, gz)
    assert text == data and int(off[-1]) == len(data) and int(off[-2]) == len(data)
    # a plain gzip file is one member; two concatenated plain members are ONE unit for the splitter, which the device
    # rejects (it cannot size the output from the last trailer alone) — the host layer keeps zlib for those
    plain = gzip.compress(data[:100_000], mtime=0)
    assert list(ffi.gzip_members(plain)) == [0, len(plain)]
    both = plain + gzip.compress(b"tail\n", mtime=0)
    assert list(ffi.gzip_members(both)) == [0, len(both)]
    with pytest.raises(ffi.GtarsGpuError):
        ffi.gunzip(ctx, both)
    assert list(ffi.gzip_members(b"")) == [0]


def test_corrupt_members_are_errors(ctx):
    from gtars_b200 import ffi
    data = _texts()["bed"]
    good = gzip.compress(data, mtime=0)
    n = len(good)
    cases = {"crc": good[:n - 8] + bytes([good[n - 8] ^ 1]) + good[n - 7:],
             "isize": good[:n - 4] + struct.pack("<I", len(data) + 1),
             "magic": b"\x1f\x8c" + good[2:],
             "truncated": good[: n // 2] + good[n - 8:],
             "flipped_payload": good[:n // 3] + bytes([good[n // 3] ^ 0x55]) + good[n // 3 + 1:]}
    for name, blob in cases.items():
        with pytest.raises(ffi.GtarsGpuError):
            ffi.gunzip(ctx, blob, np.array([0, len(blob)], dtype=np.uint64))
    # one bad member among good ones is named
    gz = good + cases["crc"] + good
    mo = np.array([0, n, 2 * n, 3 * n], dtype=np.uint64)
    with pytest.raises(ffi.GtarsGpuError, match="member 1"):
        ffi.gunzip(ctx, gz, mo)
    text, _ = ffi.gunzip(ctx, good + good, np.array([0, n, 2 * n], dtype=np.uint64))
    assert text == data + data


def test_text_goes_from_the_inflater_into_the_parser(ctx, tmp_path):
    """gtgpu_tokenize_bed_gz / gtgpu_tokenize_fragments_gz == their text forms; the host layer picks the device inflater for
    bgzip'ed files."""
    from gtars_b200 import api, ffi, synth
    u = synth.make_universe(30_000)
    offs = u["chrom_offsets"].numpy().astype(np.uint64)
    s, e, v = (u[k].numpy().view(np.uint32) for k in ("g_start", "g_end", "g_val"))
    ix = ffi.Index(ctx, ffi.KIND_BITS, offs, s, e, v)
    names = list(synth.CHROM_NAMES)
    q = synth.make_query_files(u, 1, 120_000, unknown_frac_ppm=300)
    qc, qs, qe = (q[k].numpy() for k in ("chr", "start", "end"))
    name_arr = np.array(names + ["chrUn"])
    bed = ("\n".join(map("\t".join, zip(name_arr[qc], qs.astype(str), qe.astype(str)))) + "\n").encode()
    want = ix.tokenize_bed(bed, names, int(u["unk_id"]))
    gz = bgzf(bed)
    assert np.array_equal(ix.tokenize_bed_gz(gz, names, int(u["unk_id"])), want)
    one = gzip.compress(bed, mtime=0)                                      # a single member works too (one warp)
    assert np.array_equal(ix.tokenize_bed_gz(one, names, int(u["unk_id"])), want)
    # fragments: chr start end barcode count
    rng = np.random.default_rng(3)
    bcs = np.array([f"BC{k:05d}-1" for k in range(700)])
    frag = ("# comment\n" + "\n".join(map("\t".join, zip(name_arr[qc], qs.astype(str), qe.astype(str), bcs[rng.integers(0, 700, len(qc))],
                                                          rng.integers(1, 5, len(qc)).astype(str)))) + "\n").encode()
    b1, o1, i1 = ix.tokenize_fragments_text(frag, names, int(u["unk_id"]))
    fgz = bgzf(frag)
    b2, o2, i2 = ix.tokenize_fragments_text(fgz, names, int(u["unk_id"]), gz_member_offsets=ffi.gzip_members(fgz))
    assert b1 == b2 and np.array_equal(o1, o2) and np.array_equal(i1, i2)
    ix.close()
    # host layer: a bgzip'ed universe-sized query file through Tokenizer.encode_bed_file
    upath, qpath = str(tmp_path / "universe.bed"), str(tmp_path / "query.bed.gz")
    uc, us, ue = (u[k].numpy() for k in ("chr", "start", "end"))
    with open(upath, "w") as f:
        f.write("\n".join(map("\t".join, zip(np.array(names)[uc], us.astype(str), ue.astype(str)))) + "\n")
    big = bed * 10                                                          # >= 16 members: the device path
    with open(qpath, "wb") as f:
        f.write(bgzf(big))
    tok = api.Tokenizer(upath)
    plain = str(tmp_path / "query.bed")
    with open(plain, "wb") as f:
        f.write(big)
    assert tok.encode_bed_file(qpath) == tok.encode_bed_file(plain)
