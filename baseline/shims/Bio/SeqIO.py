"""Minimal `Bio.SeqIO.parse` for FASTA/FASTQ (lengths only matter to the reference)."""


class _Rec:
    def __init__(self, rid, seq, qual=None):
        self.id = rid
        self.seq = seq
        self.letter_annotations = {} if qual is None else {"phred_quality": [ord(c) - 33 for c in qual]}


def parse(handle, fmt):
    if fmt == "fasta":
        rid, chunks = None, []
        for line in handle:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if rid is not None:
                    yield _Rec(rid, "".join(chunks))
                rid, chunks = line[1:].split()[0] if len(line) > 1 else "", []
            elif rid is not None:
                chunks.append(line.strip())
        if rid is not None:
            yield _Rec(rid, "".join(chunks))
    elif fmt == "fastq":
        while True:
            head = handle.readline()
            if not head:
                return
            seq = handle.readline().rstrip("\n")
            handle.readline()
            qual = handle.readline().rstrip("\n")
            yield _Rec(head[1:].split()[0] if len(head) > 1 else "", seq, qual)
    else:
        raise ValueError(fmt)
