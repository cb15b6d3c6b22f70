"""Contig coordinates of the concatenated genome (mirrors src/esloco/fasta_operations.py).  Only contig names
and LENGTHS are needed by the whole pipeline, so the FASTA is streamed and never held in memory (the reference
joins all sequences into one string, fasta_operations.py:17-22)."""
import gzip


def fasta_contig_lengths(path_to_fasta, chunk_bytes=1 << 24):
    """Yield (record id, sequence length) in file order; id = first whitespace-delimited token of the header.
    Byte-level scan in 16 MB chunks (find the next ">" / newline, count everything else that is not a line break),
    so an hg38-sized FASTA is measured at disk speed instead of through tens of millions of Python line objects."""
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
