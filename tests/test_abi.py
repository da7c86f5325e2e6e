"""CPU-side checks of the drop-in boundary: libgtars_gpu.so loads and exports every symbol the header declares,
the ctypes binding covers exactly that set, and the library refuses to run without a CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gtars_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gtgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from gtars_b200 import ffi
    syms = _declared_symbols()
    assert len(syms) >= 20
    L = ctypes.CDLL(ffi.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/gtars_gpu.h but not exported"


def test_binding_matches_header():
    from gtars_b200 import ffi
    assert sorted(ffi.SIGNATURES) == _declared_symbols()
    ffi.lib()


def test_no_cpu_fallback():
    """Without a GPU, init must fail loudly (GTGPU_ERR_CUDA) rather than fall back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from gtars_b200 import ffi
    with pytest.raises(ffi.GtarsGpuError) as ei:
        ffi.Context(0)
    assert "no CPU fallback" in str(ei.value) or ei.value.code == 2


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under gtars_b200/ may reference it."""
    pkg = os.path.join(ROOT, "gtars_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), f"{os.path.join(dirpath, f)} mentions the oracle"


def test_gtok_files_match_reference_bytes(golden, tmp_path):
    """K13: the .gtok files under the reference's tests/data/out, byte for byte (gtars-io/src/gtok.rs:126-300).
    Host-only code: runs without a GPU."""
    from gtars_b200 import api
    k = golden[1]["K13_gtok"]
    for name, want in k["tokens"].items():
        raw = bytes.fromhex(k["files"][name]["hex"])
        p = str(tmp_path / name)
        api.write_tokens_to_gtok(p, want)
        assert open(p, "rb").read() == raw, name
        open(p, "wb").write(raw)
        assert api.read_tokens_from_gtok(p) == want
    big = str(tmp_path / "big.gtok")
    api.write_tokens_to_gtok(big, [1, 70000, 3])                     # one token above u16 -> u32 flag
    assert open(big, "rb").read()[:5] == b"GTOK\x02" and api.read_tokens_from_gtok(big) == [1, 70000, 3]
    api.init_gtok_file(big)
    assert open(big, "rb").read() == b"GTOK\x02"
    api.append_tokens_to_gtok_file(big, [5, 0x12345678])
    assert api.read_tokens_from_gtok(big) == [5, 0x12345678]
    small = str(tmp_path / "small.gtok")
    api.write_tokens_to_gtok(small, [7])
    api.append_tokens_to_gtok_file(small, [65537])                  # u16 file: appended tokens are truncated (gtok.rs:278-284)
    assert api.read_tokens_from_gtok(small) == [7, 1]
    bad = str(tmp_path / "bad.gtok")
    open(bad, "wb").write(b"NOPE\x01\x00\x00")
    with pytest.raises(api.GtarsError):
        api.read_tokens_from_gtok(bad)


def test_host_regionset_parser_matches_oracle(tmp_path, fixture_dir):
    """The host layer's RegionSet::try_from (one pass over the bytes, integer-key stable sort) against the oracle's
    line-by-line restatement: fixture files, a large shuffled file with many ties, and fuzzed / malformed texts.
    Host-only code: runs without a GPU."""
    import numpy as np
    from gtars_b200 import api
    from oracle import oracle as orc

    def same(p):
        try:
            ref = orc.regionset_from_file(p)
        except ValueError:
            ref = None
        try:
            got = [(r.chr, r.start, r.end) for r in api.RegionSet(p)]
        except api.GtarsError:
            got = None
        assert (ref is None) == (got is None), p
        if ref is not None:
            assert got == [tuple(r) for r in ref], p
        return ref is not None

    for rel in ("to_tokenize.bed", "tokenizers/peaks.bed", "tokenizers/peaks.bed.gz", "tokenizers/peaks.scored.bed",
                "igd_file_list_01/igd_bed_file_1.bed", "consensus/consensus1.bed"):
        assert same(os.path.join(fixture_dir, rel))
    rng = np.random.default_rng(3)
    names = ["chr1", "chr10", "chr2", "chrX", "chr1_alt"]
    p = str(tmp_path / "ties.bed")
    with open(p, "w") as f:
        f.write("\n".join(f"{names[int(a)]}\t{int(b)}\t{int(b) + i % 7}\tn{i}" for i, (a, b) in
                          enumerate(zip(rng.integers(0, 5, 50_000), rng.integers(0, 2_000, 50_000)))) + "\n")
    assert same(p)
    rs = api.RegionSet(p)
    assert len(rs) == 50_000
    starts = ["0", "5", "+7", "12", "4294967295", "4294967296", "-1", "", " 3", "3 ", "007"]
    tails = ["", "\tname", "\tname\t0\t+", "\t", "\t\t", "\ta b"]
    n_ok = n_bad = 0
    for it in range(300):
        lines = []
        for _ in range(int(rng.integers(1, 7))):
            k = rng.random()
            if k < 0.08:
                lines.append(["#c", "track t", "browser b"][int(rng.integers(0, 3))])
            elif k < 0.12:
                lines.append(["", "chr1", "chr1\t5", "chr1 5 9"][int(rng.integers(0, 4))])
            else:
                a = starts[int(rng.integers(0, len(starts)))] if rng.random() < 0.15 else str(int(rng.integers(0, 1000)))
                b = starts[int(rng.integers(0, len(starts)))] if rng.random() < 0.15 else str(int(rng.integers(0, 2000)))
                lines.append(names[int(rng.integers(0, 5))] + "\t" + a + "\t" + b + tails[int(rng.integers(0, len(tails)))])
        if rng.random() < 0.1:
            lines.insert(0, "chrom\tstart\tend")
        nl = "\r\n" if rng.random() < 0.2 else "\n"
        q = str(tmp_path / f"f{it}.bed")
        with open(q, "w", newline="") as f:
            f.write(nl.join(lines) + (nl if rng.random() < 0.7 else ""))
        if same(q):
            n_ok += 1
        else:
            n_bad += 1
    assert n_ok > 50 and n_bad > 50


def test_host_igd_writer_matches_oracle_bytes(golden, fixture_dir, tmp_path):
    """Igd::save (igd.rs:418-486) as written by the host layer == the oracle's restatement, byte for byte, for the
    reference's LOLA / IGD fixture databases; the .tsv carries index, name, region count and the two-decimal average."""
    from gtars_b200 import api
    from oracle import oracle as orc
    k = golden[1]["K8_inputs"]
    for db, rels in k["dbs"].items():
        sets = [api.RegionSet(os.path.join(fixture_dir, r)) for r in rels]
        names = [f"{i}_{os.path.basename(r)}" for i, r in enumerate(rels)]
        p = str(tmp_path / f"{db}.igd")
        api.write_igd_file(sets, names, p)
        chrom_ids, o = {}, orc.Igd()
        for f, rs in enumerate(sets):
            for r in rs:
                if r.start < r.end:
                    o.add(chrom_ids.setdefault(r.chr, len(chrom_ids)), r.start, r.end, 0, f)
        o.finalize()
        po = str(tmp_path / f"{db}.oracle.igd")
        o.save(po, list(chrom_ids))
        assert open(p, "rb").read() == open(po, "rb").read(), db
        tsv = open(str(tmp_path / f"{db}.tsv")).read().splitlines()
        assert tsv[0] == "Index\tFile\tNumber of Regions\tAvg size" and len(tsv) == 1 + len(sets)
        for i, rs in enumerate(sets):
            kept = [r for r in rs if r.start < r.end]
            assert tsv[1 + i] == "%d\t%s\t%d\t%.2f" % (i, names[i], len(kept), sum(r.end - r.start for r in kept) / len(kept))


def test_device_code_is_sm100a_and_keeps_its_resource_budgets():
    """The shipped library embeds sm_100a code only, contains the kernels the design names, and the kernels whose
    occupancy the measured numbers rest on keep their register / stack budgets (a silent spill costs a CTA per SM)."""
    import shutil
    import subprocess
    from gtars_b200 import ffi
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    elfs = subprocess.run([tool, "-lelf", ffi.LIB_PATH], capture_output=True, text=True, check=True).stdout.split("\n")
    cubins = [l.split()[-1] for l in elfs if l.strip().startswith("ELF file")]
    assert cubins and all(".sm_100a." in c for c in cubins), cubins
    usage = subprocess.run([tool, "-res-usage", ffi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)", usage):
        res[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
    for name in ("fused_find_kernel", "count_kernel_x4", "count_stage_kernel", "count_runs_kernel",
                 "count_unstage_kernel", "igd_count_kernel", "radix_scatter_kernel", "scan_down_kernel", "score_hist_kernel",
                 "ingest_parse_lines_kernel", "untranspose_blocks_kernel", "radix_hist_group_kernel", "gunzip_kernel"):
        assert any(name in k for k in res), f"{name} missing from the device code"
    # radix scatter, default shape (512 threads x 16 rounds, two blocks per SM = 64 registers), all three modes: no stack
    scat = [v for k, v in res.items() if "radix_scatter_kernelILi" in k and "ELi512ELi16ELi2E" in k]
    assert len(scat) >= 4 and all(reg <= 64 and stack == 0 for reg, stack, _ in scat), scat
    # lean fused find kernel: template flags <ROWS=4, DESC, FILTER, OFFS, LEAN=1>; no stack; 40 registers = 6 CTAs per SM
    # (tokenize), 48 registers = 5 CTAs per SM when it also writes per-query offsets (find, fragments)
    # (flags after LEAN: UNK1, TAG — the tagged variant of fragment tokenization has the 48-register budget too)
    lean = [(re.match(r"(?:Lb[01]E){2}Lb([01])ELb1ELb[01]ELb([01])E", k.split("fused_find_kernelILi4E")[1]), v) for k, v in res.items()
            if "fused_find_kernelILi4E" in k]
    lean = [(m.group(1) == "1" or m.group(2) == "1", v) for m, v in lean if m]
    assert lean and any(o for o, _ in lean) and any(not o for o, _ in lean), "lean instantiations of the fused kernel missing"
    for offs, (reg, stack, shared) in lean:
        assert reg <= (48 if offs else 40) and stack == 0 and shared <= 24 * 1024, (offs, reg, stack, shared)
    part = [v for k, v in res.items() if "count_stage_kernel" in k or "count_unstage_kernel" in k]
    assert part and all(reg <= 32 and stack <= 32 for reg, stack, _ in part), part   # two 1024-thread CTAs per SM


def test_rust_sys_stub_declares_every_entry_point():
    """bindings/rust/gtars-overlaprs-sys (what a gtars maintainer would vendor; cannot be compiled here) stays in step
    with the header: every exported entry point has an `extern "C"` declaration."""
    rs = open(os.path.join(ROOT, "bindings", "rust", "gtars-overlaprs-sys", "src", "lib.rs")).read()
    missing = [s for s in _declared_symbols() if not re.search(r"\bfn\s+" + s + r"\b", rs)]
    assert not missing, missing
    # build.rs compiles exactly the sources the Makefile links into libgtars_gpu.so (a missing one is a link failure the
    # day someone runs cargo)
    mk = open(os.path.join(ROOT, "gtars_b200", "csrc", "Makefile")).read()
    srcs = sorted(re.findall(r"cuda/(\w+)\.cu", re.search(r"^SRCS := (.*)$", mk, re.M).group(1)))
    brs = open(os.path.join(ROOT, "bindings", "rust", "gtars-overlaprs-sys", "build.rs")).read()
    listed = sorted(re.findall(r'"(\w+)"', re.search(r"for f in \[(.*?)\]", brs, re.S).group(1)))
    assert listed == srcs, (listed, srcs)
