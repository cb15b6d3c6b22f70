"""Host set-up (FASTA / BED / insertion generator / mask tree) against what the unmodified reference produced
for the same files and numpy seed (golden fixtures)."""
import os

import numpy as np
import pytest

from golden_util import Golden, case_names

CASES = [c for c in case_names() if Golden(c).meta.get("contigs")]


def write_fasta(path, contigs, width=70):
    with open(path, "w") as fh:
        for name, n in contigs:
            fh.write(f">{name} some description\n")
            line = "ACGT" * 25000
            full, rest = divmod(n, len(line))
            for _ in range(full):
                fh.write(line + "\n")
            if rest:
                fh.write(line[:rest] + "\n")


def write_bed(path, header, rows):
    with open(path, "w") as fh:
        if header:
            fh.write("\t".join(header) + "\n")
        for r in rows:
            fh.write("\t".join(str(x) for x in r) + "\n")


@pytest.mark.parametrize("name", CASES)
def test_setup_matches_reference(name, tmp_path):
    from esloco_b200.genome_generation import create_barcoded_insertion_genome, create_barcoded_roi_genome
    from esloco_b200.utils import build_mask_tree, mask_tree_to_arrays, targets_to_arrays
    g = Golden(name)
    m = g.meta
    fasta = str(tmp_path / "ref.fa")
    write_fasta(fasta, m["contigs"])
    blocked = None
    if m.get("blocked_rows"):
        blocked = str(tmp_path / "blocked.bed")
        write_bed(blocked, m["blocked_header"], m["blocked_rows"])
    np.random.seed(m["setup_seed"])
    if m["mode"] == "I":
        G, targets, masked, cdir = create_barcoded_insertion_genome(
            1, fasta, None, blocked, None, m["insertion_length"], m["insertion_numbers"], None, m["n_barcodes"])
    else:
        roi = str(tmp_path / "roi.bed")
        write_bed(roi, ["chrom", "start", "end", "id"], m["roi_rows"])
        targets, _bed, masked, G, cdir = create_barcoded_roi_genome(fasta, None, roi, m["n_barcodes"], blocked)
    assert G == m["genome_size"]
    assert {k: list(v) for k, v in cdir.items()} == m["chromosome_dir"]
    keys, ts, te, tb = targets_to_arrays(targets, m["n_barcodes"])
    assert keys == m["target_keys"]
    assert np.array_equal(ts, g["target_start"]) and np.array_equal(te, g["target_end"])
    tree = build_mask_tree(masked) if masked else None
    ms, me, mw, mb = mask_tree_to_arrays(tree, m["n_barcodes"])
    order = np.lexsort((mb, me, ms)); gorder = np.lexsort((g["mask_barcode"], g["mask_end"], g["mask_start"]))
    assert np.array_equal(ms[order], g["mask_start"][gorder]) and np.array_equal(me[order], g["mask_end"][gorder])
    assert np.array_equal(mw[order], g["mask_weight"][gorder]) and np.array_equal(mb[order], g["mask_barcode"][gorder])


def test_fasta_filters_and_unrestricted(tmp_path):
    from esloco_b200.fasta_operations import pseudo_fasta_coordinates
    p = str(tmp_path / "x.fa")
    write_fasta(p, [("chr1", 1000), ("chr1_alt", 500), ("chrM", 300), ("chr2", 700)])
    assert pseudo_fasta_coordinates(p) == (1700, {"chr1": [0, 1000], "chr2": [1000, 1700]})
    total, d = pseudo_fasta_coordinates(p, "unrestricted")
    assert total == 2500 and list(d) == ["chr1", "chr1-alt", "chrM", "chr2"]


def test_seq_read_data(tmp_path):
    from esloco_b200.config_handler import seq_read_data
    fq = tmp_path / "r.fastq"
    lens = [10, 250, 3000, 5, 777]
    fq.write_text("".join(f"@r{i}\n{'A' * n}\n+\n{'I' * n}\n" for i, n in enumerate(lens)))
    assert seq_read_data(str(fq), distribution=True, min_read_length=9).tolist() == [10, 250, 3000, 777]
    assert seq_read_data(str(fq), min_read_length=0) == np.mean(lens)
    fa = tmp_path / "r.fa"
    fa.write_text(">a\nACGT\nAC\n>b\nA\n")
    assert seq_read_data(str(fa), distribution=True).tolist() == [6, 1]
    (tmp_path / "r.txt").write_text("x\n")
    with pytest.raises(ValueError):
        seq_read_data(str(tmp_path / "r.txt"))


def test_weighted_probabilities_and_mask_keys():
    from esloco_b200.read_operations import barcode_probabilities, get_weighted_probabilities
    from esloco_b200.utils import build_mask_tree, mask_tree_to_arrays
    from oracle import oracle as O
    for B, w in [(4, {"0": 10}), (4, None), (3, {"1": 2}), (24, {"0": 10, "1": 5}), (2, {"Barcode": 3})]:
        assert np.array_equal(barcode_probabilities(B, w), O.barcode_probabilities(B, w))
    assert get_weighted_probabilities("Barcode_1", 4, {"0": 10}) == 4 / ((10 + 4 - 1) * 4)
    tree = build_mask_tree({"a": {"start": 5, "end": 9, "weight": "0.5", "barcode": '["0","1"]'},
                            "b": {"start": 1, "end": 2, "weight": 1, "barcode": []},
                            "c": {"start": 3, "end": 4, "weight": 1, "barcode": [0]}})  # int key: unreachable in the reference
    assert set(tree) == {"0", "1", "ALL", 0}
    ms, me, mw, mb = mask_tree_to_arrays(tree, 2)
    assert sorted(zip(ms, me, mw, mb)) == [(1, 2, 1.0, -1), (5, 9, 0.5, 0), (5, 9, 0.5, 1)]


def test_native_scanner_matches_python_pass(tmp_path):
    """esl_scan_read_lengths (C++, one streaming pass, gzip transparent) == the pure-Python pass, FASTA and FASTQ,
    multi-line FASTA records, CRLF line ends, blank lines, min_read_length filter, malformed FASTQ."""
    import gzip
    from esloco_b200 import config_handler as ch
    rng = np.random.RandomState(4)
    lens = rng.randint(1, 5000, 300).tolist()
    fq = tmp_path / "r.fastq"
    fq.write_text("".join(f"@r{i} extra\n{'A' * n}\n+\n{'I' * n}\n" for i, n in enumerate(lens)))
    fa = tmp_path / "r.fa"
    fa.write_text("".join(f">c{i}\n" + "\n".join("ACGT"[:1] * 0 + "C" * min(70, n - k) for k in range(0, n, 70)) + "\r\n" for i, n in enumerate(lens)))
    gz = tmp_path / "r.fq.gz"
    with gzip.open(gz, "wt") as fh:
        fh.write(fq.read_text())
    for path in (fq, fa, gz):
        for m in (0, 1000):
            native = ch._scan_native(str(path), str(path).endswith(ch.FASTQ_EXTS), m)
            assert native is not None
            assert native.tolist() == [n for n in lens if n > m], path
            assert ch.seq_read_data(str(path), distribution=True, min_read_length=m).tolist() == native.tolist()
    assert ch.seq_read_data(str(fq)) == np.mean(lens)
    bad = tmp_path / "bad.fastq"
    bad.write_text("@r0\nACGT\n+\n\n")
    with pytest.raises(ValueError):
        ch.seq_read_data(str(bad))
    with pytest.raises(ValueError):
        ch.seq_read_data(str(fq), min_read_length=10**7)


def test_fasta_byte_scanner_chunk_boundaries(tmp_path):
    """The chunked byte-level FASTA scan gives the same contig lengths for any chunk size (headers and sequence lines
    split across chunk boundaries, CRLF, blank lines, gzip, '>' inside a description)."""
    import gzip
    from esloco_b200.fasta_operations import fasta_contig_lengths
    text = ">chr1 desc >weird\nACGT\nAC\r\n\n>chr2\n" + "G" * 205 + "\n" + "T" * 17 + "\n>empty\n>chr3 last"
    p = tmp_path / "a.fa"
    p.write_text(text)
    expect = [("chr1", 6), ("chr2", 222), ("empty", 0), ("chr3", 0)]
    for cb in (1, 2, 3, 7, 64, 1 << 20):
        assert list(fasta_contig_lengths(str(p), cb)) == expect, cb
    gz = tmp_path / "a.fa.gz"
    with gzip.open(gz, "wt") as fh:
        fh.write(text)
    assert list(fasta_contig_lengths(str(gz), 5)) == expect


def test_parse_config_matches_reference(tmp_path):
    """parse_config returns the reference's parameter dictionary (keys, types, defaults, derived values) for the INI
    files recorded from the unmodified reference in tests/golden/config_golden.json (+ the `overlap_lists`, `length_sampler` and `label_columns` extensions)."""
    import json
    from esloco_b200.config_handler import parse_config
    tmp = str(tmp_path)
    with open(os.path.join(tmp, "reads.fastq"), "w") as fh:
        for i, n in enumerate((10, 30, 55, 100, 21)):
            fh.write(f"@r{i}\n{'A' * n}\n+\n{'I' * n}\n")
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_golden.json")))
    assert len(cases) >= 4
    for k, case in enumerate(cases):
        ini = os.path.join(tmp, f"c{k}.ini")
        open(ini, "w").write(case["ini"].replace("{tmp}", tmp))
        got = parse_config(ini)
        assert got.pop("overlap_lists") is False
        assert got.pop("length_sampler") == "quantile"
        assert got.pop("label_columns") == "barcode"
        got = json.loads(json.dumps(got, default=str).replace(tmp, "{tmp}"))
        assert got == case["params"], k
    bad = os.path.join(tmp, "bad.ini")
    open(bad, "w").write("[COMMON]\nmode = X\nreference_genome_path = a\n")
    with pytest.raises(SystemExit):
        parse_config(bad)


def _jsonable(x):
    if isinstance(x, dict):
        return {str(k): _jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, np.integer):
        return int(x)
    if isinstance(x, np.floating):
        return float(x)
    return x


def test_bed_variants_match_reference(tmp_path):
    """Header-less BEDs, BED-guided insertion placement (length-weighted, explicit weights, one insertion per row) and
    Poisson insertion counts: same targets / masked regions as the unmodified reference for the same numpy seed, and
    the numpy stream is left at the same position (tests/golden/setup_golden.json)."""
    import json
    from esloco_b200.genome_generation import create_barcoded_insertion_genome, create_barcoded_roi_genome
    cases = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "setup_golden.json")))
    assert len(cases) >= 5
    for g in cases:
        c = g["case"]
        tmp = str(tmp_path / c["name"])
        os.makedirs(tmp)
        fasta = os.path.join(tmp, "g.fa")
        write_fasta(fasta, g["contigs"])

        def bed(spec, name):
            if spec is None:
                return None
            p = os.path.join(tmp, name)
            write_bed(p, spec[0], spec[1])
            return p
        blocked = bed(c.get("blocked"), "blocked.bed")
        np.random.seed(c["seed"])
        if c["mode"] == "I":
            G, targets, masked, cdir = create_barcoded_insertion_genome(
                1, fasta, bed(c.get("guide"), "guide.bed"), blocked, None, c["insertion_length"], c["insertion_numbers"], c["dist"],
                c["n_barcodes"])
        else:
            targets, _bed, masked, G, cdir = create_barcoded_roi_genome(fasta, None, bed(c["roi"], "roi.bed"), c["n_barcodes"], blocked)
        assert G == g["genome_size"], c["name"]
        assert _jsonable(cdir) == g["chromosome_dir"], c["name"]
        assert _jsonable(targets) == g["targets"], c["name"]
        assert _jsonable(masked) == g["masked"], c["name"]
        assert float(np.random.rand()) == g["stream_tail"], c["name"]


def test_targets_to_arrays_shapes_and_barcode_rule():
    """The column-wise fast path (I mode: every value a [start, end] list) and the per-entry path (ROI dicts, mixed or
    ragged values) give the same arrays; the barcode of a key is its second `_` token iff that is the canonical decimal
    spelling of a barcode number below n_barcodes (counting.py:37 compares strings)."""
    from esloco_b200.utils import targets_to_arrays
    ins = {"Barcode_0_insertion_0": [10, 20], "Barcode_1_insertion_0": [30, 45], "Barcode_01_insertion_1": [5, 6],
           "Barcode_7_insertion_2": [1, 2], "nounderscore": [3, 4], "x_-1_y": [8, 9]}
    keys, s, e, b = targets_to_arrays(ins, 2)
    assert keys == list(ins) and s.tolist() == [10, 30, 5, 1, 3, 8] and e.tolist() == [20, 45, 6, 2, 4, 9]
    assert b.tolist() == [0, 1, -1, -1, -1, -1] and s.dtype == np.int64 and b.dtype == np.int32
    roi = {"R0_0": {"start": 10, "end": 20, "weight": 1}, "R0_1": {"start": 10, "end": 20, "weight": 1}, "R1_0": [7, 9, "extra"]}
    keys, s, e, b = targets_to_arrays(roi, 2)
    assert s.tolist() == [10, 10, 7] and e.tolist() == [20, 20, 9] and b.tolist() == [0, 1, 0]
    big = {f"Barcode_{i % 3}_insertion_{i}": [i * 10, i * 10 + 5] for i in range(5000)}
    keys, s, e, b = targets_to_arrays(big, 3)
    assert (s == np.arange(5000) * 10).all() and (e - s == 5).all() and (b == np.arange(5000) % 3).all()
    np_vals = {"a_0": [np.int64(4), np.int64(9)], "b_1": [np.int64(1), 2]}
    keys, s, e, b = targets_to_arrays(np_vals, 2)
    assert s.tolist() == [4, 1] and e.tolist() == [9, 2]
    assert targets_to_arrays({}, 4)[1].size == 0


def test_native_contig_scanner_matches_python_scan(tmp_path):
    """esl_scan_fasta_contigs (C++, one pass, SSE2 block counting, gzip transparent) == the chunked Python scan: ids and
    lengths for CRLF / blank lines / blanks inside sequence lines / '>' inside a description / a header at the end of
    the file / text before the first header / an empty file / more records and id bytes than the first call's buffers
    hold (the retry path), plain and gzip."""
    import gzip
    from esloco_b200 import fasta_operations as fo
    rng = np.random.RandomState(3)
    texts = {
        "tricky": ">chr1 desc >weird\nACGT\nAC\r\n\n>chr2\n" + "G" * 205 + "\n" + "T" * 17 + "\n>empty\n>chr3 last",
        "junk": "text before\n>  \n A C\tG \n>x\n",
        "empty": "",
        "noheader": "ACGT\nACGT\n",
        "many": "".join(f">scaffold_{i}_{'n' * (i % 40)} d\n" + "\n".join("A" * w for w in rng.randint(1, 90, rng.randint(0, 5))) + "\n" for i in range(6000)),
        "long": ">a\n" + ("ACGT" * 20 + "\n") * 200_000 + ">b x\n" + "N" * 3_000_001,
    }
    for name, text in texts.items():
        p = tmp_path / f"{name}.fa"
        p.write_bytes(text.encode())
        want = list(fo.fasta_contig_lengths(str(p), 1 << 16))
        assert fo._contigs_native(str(p)) == want, name
        assert list(fo.fasta_contig_lengths(str(p))) == want, name
        gz = tmp_path / f"{name}.fa.gz"
        with gzip.open(gz, "wb") as fh:
            fh.write(text.encode())
        assert fo._contigs_native(str(gz)) == want, name
    assert dict(fo.fasta_contig_lengths(str(tmp_path / "long.fa"))) == {"a": 16_000_000, "b": 3_000_001}
    with pytest.raises(OSError):
        fo._contigs_native(str(tmp_path / "missing.fa"))


def test_shift_insertions_equals_the_sequential_walk():
    """esl_shift_insertions against the reference's own loop written out (create_insertion_genome.py:60-64,76-80: every
    new insertion pushes the earlier ones that start at or after it), random positions with many ties, 0..600
    insertions, several lengths."""
    from esloco_b200.create_insertion_genome import _shifted_starts
    rng = np.random.RandomState(8)
    for trial in range(60):
        n = int(rng.choice([0, 1, 2, 17, 200, 600]))
        length = int(rng.choice([1, 500, 5000, 10**7]))
        span = int(rng.choice([5, 1000, 10**9]))  # a span of 5 makes nearly every position a tie
        pos = rng.randint(0, span, n).astype(np.int64)
        walk = []
        for p in pos.tolist():
            for k in range(len(walk)):
                if walk[k] >= p:
                    walk[k] += length
            walk.append(p)
        assert _shifted_starts(pos, length).tolist() == walk, (trial, n, length, span)


def test_contig_scanner_random_files(tmp_path):
    """Random FASTA text (line widths 1-120, LF / CRLF, blank lines, blanks and tabs inside sequence lines, empty records,
    descriptions, leading junk, with and without a final line break) through the native scanner and the Python scan."""
    from esloco_b200 import fasta_operations as fo
    rng = np.random.RandomState(21)
    for trial in range(40):
        parts = []
        if rng.rand() < 0.3:
            parts.append("junk line\n")
        for r in range(int(rng.choice([0, 1, 3, 40]))):
            nl = "\r\n" if rng.rand() < 0.3 else "\n"
            parts.append(">" + ("" if rng.rand() < 0.05 else f"c{trial}_{r}") + (" some text > here" if rng.rand() < 0.4 else "") + nl)
            for _ in range(int(rng.choice([0, 1, 5, 60]))):
                line = "".join(rng.choice(list("ACGTN"), int(rng.randint(1, 121))))
                if rng.rand() < 0.1:
                    line = line[: len(line) // 2] + rng.choice([" ", "\t", "  "]) + line[len(line) // 2:]
                parts.append(line + nl)
                if rng.rand() < 0.05:
                    parts.append(nl)
        text = "".join(parts)
        if text.endswith("\n") and rng.rand() < 0.5:
            text = text[:-1].rstrip("\r")
        p = tmp_path / f"r{trial}.fa"
        p.write_bytes(text.encode())
        want = list(fo.fasta_contig_lengths(str(p), int(rng.choice([3, 64, 1 << 16]))))
        assert fo._contigs_native(str(p)) == want, trial


def test_read_length_scan_second_pass_when_the_guess_is_short(tmp_path):
    """The native read-length scan sizes its output from the file size (one pass for long reads); a file with more
    records than that guess (70 k reads of 3-9 bases) takes the second pass and still returns every length."""
    from esloco_b200 import config_handler as ch
    rng = np.random.RandomState(2)
    lens = rng.randint(3, 10, 70_000)
    p = tmp_path / "short.fa"
    p.write_text("".join(f">r{i}\n{'A' * n}\n" for i, n in enumerate(lens)))
    got = ch._scan_native(str(p), False, 4)
    assert got is not None and got.tolist() == [int(n) for n in lens if n > 4]
