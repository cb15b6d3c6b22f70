"""Contig coordinates of the concatenated genome (mirrors src/esloco/fasta_operations.py).  Only contig names
and LENGTHS are needed by the whole pipeline, so the FASTA is streamed and never held in memory (the reference
joins all sequences into one string, fasta_operations.py:17-22)."""
import gzip


def _contigs_native(path_to_fasta):
    """[(id, length)] through the native one-pass scanner (esl_scan_fasta_contigs); None when the library is not built."""
    import ctypes as C

    import numpy as np

    from . import _native
    try:
        lib = _native.load()
    except _native.NativeLibraryError:  # library not built yet: the Python scan below does the same job, slower
        return None
    names_cap, max_records = 1 << 16, 1 << 12
    while True:
        names = C.create_string_buffer(names_cap)
        lengths = np.zeros(max_records, np.int64)
        need = C.c_int64()
        n = lib.esl_scan_fasta_contigs(str(path_to_fasta).encode(), names, names_cap, lengths.ctypes.data, max_records, C.byref(need))
        if n < 0:
            raise OSError(f"cannot read {path_to_fasta}")
        if n <= max_records and need.value <= names_cap:
            ids = names.raw[:need.value].split(b"\0")[:n] if n else []
            return [(i.decode(), int(l)) for i, l in zip(ids, lengths[:n])]
        names_cap, max_records = max(names_cap, need.value + 64), max(max_records, n)  # assemblies with > 4096 scaffolds: once more, with room


def fasta_contig_lengths(path_to_fasta, chunk_bytes=None):
    """Yield (record id, sequence length) in file order; id = first whitespace-delimited token of the header.
    Default: the native one-pass scanner (csrc/seqscan.cpp; an hg38-sized FASTA at the speed of the read).  With
    chunk_bytes given, or without the built library: a byte-level scan in Python in chunks of that size (find the next
    ">" / newline, count everything else that is not a line break)."""
    if chunk_bytes is None:
        native = _contigs_native(path_to_fasta)
        if native is not None:
            yield from native
            return
        chunk_bytes = 1 << 24
    open_func = gzip.open if str(path_to_fasta).endswith(".gz") else open
    with open_func(path_to_fasta, "rb") as fh:
        name, n, header = None, 0, None  # header: bytes of a header line that is still open (None = in sequence)
        while True:
            chunk = fh.read(chunk_bytes)
            if not chunk:
                break
            pos = 0
            while pos < len(chunk):
                if header is not None:  # inside a header line: runs to the next newline, may span chunks
                    nl = chunk.find(b"\n", pos)
                    if nl < 0:
                        header += chunk[pos:]
                        pos = len(chunk)
                        break
                    header += chunk[pos:nl]
                    tokens = header.split()
                    name, n, header = (tokens[0].decode() if tokens else ""), 0, None
                    pos = nl + 1
                else:  # sequence lines up to the next record start
                    gt = chunk.find(b">", pos)
                    end = gt if gt >= 0 else len(chunk)
                    if name is not None and end > pos:
                        seg = chunk[pos:end]
                        n += len(seg) - seg.count(b"\n") - seg.count(b"\r") - seg.count(b" ") - seg.count(b"\t")
                    pos = end
                    if gt >= 0:
                        if name is not None:
                            yield name, n
                        header, pos = b"", gt + 1
        if header is not None:  # file ends inside a header line
            tokens = header.split()
            name, n = (tokens[0].decode() if tokens else ""), 0
        if name is not None:
            yield name, n


def pseudo_fasta_coordinates(path_to_fasta, restriction=None):
    """-> (total length, {contig: [global start, global end]}) (fasta_operations.py:3-22).  Contigs whose id
    contains "_" or "M" are skipped unless restriction == "unrestricted", in which case "_" becomes "-" so later
    "_"-splits stay valid."""
    entries, total = {}, 0
    for rid, n in fasta_contig_lengths(path_to_fasta):
        if restriction == "unrestricted":
            rid = rid.replace("_", "-")
        elif "_" in rid or "M" in rid:
            continue
        entries[rid] = [total, total + n]
        total += n
    return total, entries
