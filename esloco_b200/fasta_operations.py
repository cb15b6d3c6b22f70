"""Contig coordinates of the concatenated genome (mirrors src/esloco/fasta_operations.py).  Only contig names
and LENGTHS are needed by the whole pipeline, so the FASTA is streamed and never held in memory (the reference
joins all sequences into one string, fasta_operations.py:17-22)."""
import gzip


def fasta_contig_lengths(path_to_fasta):
    """Yield (record id, sequence length) in file order; id = first whitespace-delimited token of the header."""
    open_func = gzip.open if str(path_to_fasta).endswith(".gz") else open
    with open_func(path_to_fasta, "rt") as fh:
        name, n = None, 0
        for line in fh:
            if line.startswith(">"):
                if name is not None:
                    yield name, n
                head = line[1:].split()
                name, n = (head[0] if head else ""), 0
            elif name is not None:
                n += len(line.strip())
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
