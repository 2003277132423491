"""``Bio.SeqIO.parse(path, "fasta")`` as far as MetaCoAG uses it: records with ``.id``,
``.description`` and ``.seq``.  The rules are those of Bio.SeqIO.FastaIO.SimpleFastaParser: text
before the first '>' line is skipped; title = line[1:].rstrip(); id = first word of the title;
the sequence is the right-stripped lines joined, with ' ' and '\\r' removed; text-mode handle."""


class _Record:
    def __init__(self, name, description, seq):
        self.id = name
        self.description = description
        self.seq = seq


def parse(path, fmt):
    assert fmt == "fasta"
    with open(path) as fh:
        title = None
        for line in fh:
            if line[0] == ">":
                title = line[1:].rstrip()
                break
        if title is None:
            return
        lines = []
        for line in fh:
            if line[0] == ">":
                words = title.split(None, 1)
                yield _Record(words[0] if words else "", title,
                              "".join(lines).replace(" ", "").replace("\r", ""))
                lines = []
                title = line[1:].rstrip()
                continue
            lines.append(line.rstrip())
        words = title.split(None, 1)
        yield _Record(words[0] if words else "", title,
                      "".join(lines).replace(" ", "").replace("\r", ""))
