"""One record split over several GPUs (BASELINE config 4: a chromosome-sized record).

Record sharding (shard.py) cannot spread ONE record; the reference translates a record on one
worker (writer.go:57-117).  Here the record's FASTA bytes are cut into pieces at line starts, every
rank parses its piece, the ranks exchange ~200 bytes per piece (nucleotide count, the codes of the
piece's first 192 and last two nucleotides; the header length comes from piece 0), and every rank
then writes its share of the protein FASTA at its final offsets.  Output lines are never shared:
a 60-residue line (180 nucleotides) is computed by the piece that holds the last nucleotide of the
line's lowest codon, with the two nucleotides before the piece and the 192 behind it as context.
"""


def piece_ranges(buf, npieces):
    """npieces contiguous (start, end) byte ranges of buf, roughly equal, cut at line starts (just
    after a '\\n').  buf is ONE record: it starts with its header line and holds no other."""
    n = len(buf)
    first_nl = buf.find(b"\n")
    body = n if first_nl < 0 else first_nl + 1  # piece 0 keeps the whole header line
    cuts = [0]
    for k in range(1, npieces):
        pos = (n * k) // npieces
        i = buf.rfind(b"\n", 0, pos)
        c = i + 1 if i >= 0 else 0
        cuts.append(max(c, body, cuts[-1]))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(npieces)]


def plan_pieces(infos):
    """From the pieces' scan results (in piece order) to the gts_piece description of each piece."""
    total = sum(i["nucleotides"] for i in infos)
    H = infos[0]["header_len"]
    pieces = []
    before = 0
    last, second = 0, 0  # codes of the two nucleotides before the current piece (0 = the pads)
    for k, inf in enumerate(infos):
        # the 192 nucleotides behind the piece: the heads of the pieces after it, as far as needed
        nxt = b""
        for later in infos[k + 1:]:
            if len(nxt) >= 192:
                break
            nxt += bytes(later["head"])[:min(192, later["nucleotides"])]
        pieces.append(dict(nt_before=before, nt_total=total, header_len=H, carry=(last, second),
                           first=(k == 0), last=(k == len(infos) - 1), next_head=nxt[:192]))
        m = inf["nucleotides"]
        if m >= 2:
            last, second = inf["tail"]
        elif m == 1:
            last, second = inf["tail"][0], last
        before += m
    return pieces


def merge_trim(lens, results):
    """One round of the --trim walk: every piece reports (lens, resolved) for its part; exactly one
    piece holds the residue a frame's walk stands on, the others return the lengths unchanged."""
    out, done = list(lens), [False] * 6
    for f in range(6):
        settled = [r[0][f] for r in results if r[1][f]]
        if settled:
            out[f], done[f] = min(settled), True
        else:
            out[f] = min(r[0][f] for r in results)
    return out, done


def settle_trim(start_lens, walk_round, max_rounds):
    """Trimmed frame lengths of a split record (writer.go:141-163): walk_round(lens) -> [(lens, resolved)
    of every piece]; repeated until every frame's walk has stopped (one round unless a piece is
    nothing but 'X' / '*' at one of the record's ends)."""
    lens = list(start_lens)
    for _ in range(max_rounds):
        lens, done = merge_trim(lens, walk_round(lens))
        if all(done):
            return lens
    raise RuntimeError("the trim walk over the pieces did not settle")


def assemble(total, segment_lists):
    """The record's protein FASTA from the pieces' (offset, bytes) segments; checks that they
    tile [0, total) exactly."""
    out = bytearray(total)
    segs = sorted((off, data) for lst in segment_lists for off, data in lst)
    pos = 0
    for off, data in segs:
        if off != pos:
            raise ValueError(f"segments do not tile the output: gap or overlap at {pos} (next segment at {off})")
        out[off:off + len(data)] = data
        pos = off + len(data)
    if pos != total:
        raise ValueError(f"segments end at {pos}, output has {total} bytes")
    return bytes(out)


def translate_split(buf, npieces, make_translator, trim=False):
    """Single-process driver (tests, and the model of what N ranks do): one Translator per piece.
    make_translator() -> a gotranseq_b200.Translator."""
    ranges = piece_ranges(buf, npieces)
    trs = [make_translator() for _ in ranges]
    try:
        infos = [tr.piece_scan(buf[lo:hi], k == 0) for k, (tr, (lo, hi)) in enumerate(zip(trs, ranges))]
        pieces = plan_pieces(infos)
        if trim:
            lens = settle_trim(trs[0].frame_lengths(pieces[0]["nt_total"]),
                               lambda ln: [tr.piece_trim(p, ln) for tr, p in zip(trs, pieces)], len(pieces) + 2)
            for p in pieces:
                p["trim_len"] = lens
        results = [tr.piece_translate(p) for tr, p in zip(trs, pieces)]
    finally:
        for tr in trs:
            tr.close()
    total = results[0][0]
    return assemble(total, [segs for _, segs in results])


def translate_split_rank(buf, rank, world, translator, all_gather, trim=False):
    """What rank `rank` of `world` does.  all_gather(obj) -> list of every rank's obj, in rank order
    (torch.distributed.all_gather_object).  Returns (total output size, this rank's segments)."""
    lo, hi = piece_ranges(buf, world)[rank]
    info = translator.piece_scan(buf[lo:hi], rank == 0)
    infos = all_gather(info)
    piece = plan_pieces(infos)[rank]
    if trim:
        piece["trim_len"] = settle_trim(translator.frame_lengths(piece["nt_total"]),
                                        lambda ln: all_gather(translator.piece_trim(piece, ln)), world + 2)
    return translator.piece_translate(piece)
