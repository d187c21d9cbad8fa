"""Executable specification of the GPU data flow (TEST INFRASTRUCTURE, pure Python, small inputs).

This mirrors, formula for formula, what the CUDA kernels in gotranseq_b200/csrc/ compute -- it is NOT the oracle (that is oracle/transeq_oracle.c, a restatement of the Go code).
Its job is to let the index arithmetic of the kernels (symbol stream with pads, window -> frame
routing, run geometry, 60-column line jobs, tile boundaries) be checked against the oracle on the
CPU, with deliberately tiny tile sizes so that every tile-boundary case is hit.

Stages (DESIGN.md has the narrative):
  parse      raw FASTA -> dense symbol stream (codes N=0 A=1 C=2 T=3 G=4; two 0 pads in front of
             every record, two more at the end) + record table (hdr_off, hdr_end, dbase)
  size       per record: nucleotide count, frame lengths (optionally trimmed), output offsets
  residues   two formulations of the same output, both checked against the oracle:
             "windows" (round 1's first residue kernel): per tile of `tile` symbols one LUT lookup per
             symbol e for the window (e-2, e-1, e), results spread over 3 phase arrays (e mod 3) x
             {forward, reverse}; then for every (record piece, frame) a "run" whose 60-column lines
             are copied to the output;
             "lines" (k_lines, gts_lines.cu): per tile, record piece and strand the BLOCKS of 180
             nucleotides whose anchor (last symbol of the block's lowest codon) lies in the tile; a
             block is output line b of the strand's three frames, 60 codons each read straight
             from the stream
  headers    per record and frame: header bytes with _N inserted before the first space
"""

X, STAR = ord("X"), ord("*")


def encode_byte(b):
    # encodedSequence.go:25-41
    if b in b"Aa":
        return 1
    if b in b"Cc":
        return 2
    if b in b"TtUu":
        return 3
    if b in b"Gg":
        return 4
    return 0


def parse(raw):
    """Position-based state machine of the parse kernel (k_parse)."""
    n = len(raw)
    sym = [0, 0]  # slot 0's pads
    hdr_off, hdr_end, dbase = [0], [0], [2]
    LS, SEQ, HDR = 0, 1, 2
    state = LS
    for p in range(n):
        b = raw[p]
        if b == 0x0A:
            if state == HDR:
                e = p
                if raw[p - 1] == 0x0D:
                    e -= 1
                hdr_end[-1] = e
            state = LS
            continue
        if state == LS:
            if b == 0x3E:
                state = HDR
                sym += [0, 0]
                hdr_off.append(p)
                hdr_end.append(None)
                dbase.append(len(sym))
            else:
                state = SEQ
        if state == SEQ:
            # bufio.ScanLines dropCR: a CR right before the line's '\n', or as the very last byte
            if b == 0x0D and (p + 1 == n or raw[p + 1] == 0x0A):
                continue
            sym.append(encode_byte(b))
    if state == HDR:
        e = n
        if raw[n - 1] == 0x0D:
            e -= 1
        hdr_end[-1] = e
    total = len(sym)
    dbase.append(total + 2)  # sentinel
    sym += [0, 0]  # EOF pads
    return sym, hdr_off, hdr_end, dbase, total


def staden_start(n, i, alternative):
    # writer.go:64-79: rc start position of reverse frame i (0..2)
    return i if alternative else (n % 3 - i) % 3


def frame_lengths(n, alternative):
    L = []
    for k in range(3):
        L.append((n - k + 2) // 3 if n >= k else 0)
    for i in range(3):
        p = staden_start(n, i, alternative)
        L.append((n - p + 2) // 3 if n >= p else 0)
    return L


def window_aa(sym, lut, e):
    idx = sym[e - 2] + 5 * sym[e - 1] + 25 * sym[e]
    return lut[idx] & 0xFF, lut[idx] >> 8


def trimmed_lengths(sym, lut, D, n, alternative, mask):
    """k_size with trim: walk each requested frame backwards from its last residue."""
    L = frame_lengths(n, alternative)
    for f in range(6):
        if not (mask >> f) & 1:
            continue
        j = L[f]
        while j > 0:
            if f < 3:
                q = f + 3 * (j - 1)
                aa = window_aa(sym, lut, D + q + 2)[0]
            else:
                p = staden_start(n, f - 3, alternative)
                q = n - 3 - p - 3 * (j - 1)
                aa = window_aa(sym, lut, D + q + 2)[1]
            if aa != X and aa != STAR:
                break
            j -= 1
        L[f] = j
    return L


def run_size(H, Lp):
    return H + 3 + Lp + (Lp + 59) // 60


def build_lut(aa64, clean):
    """gts_codes.h build_fused_lut."""
    tcag = {1: 2, 2: 1, 3: 0, 4: 3}

    def comp(c):
        return 0 if c == 0 else ((c - 1) ^ 2) + 1

    def aa(c0, c1, c2):
        if c0 == 0 or c1 == 0:
            return X
        b0, b1 = tcag[c0], tcag[c1]
        if c2 == 0:
            q = aa64[16 * b0 + 4 * b1: 16 * b0 + 4 * b1 + 4]
            if len(set(q)) != 1:
                return X
            r = ord(q[0])
        else:
            r = ord(aa64[16 * b0 + 4 * b1 + tcag[c2]])
        if clean and r == STAR:
            return X
        return r

    lut = [0] * 125
    for c2 in range(5):
        for c1 in range(5):
            for c0 in range(5):
                lut[c0 + 5 * c1 + 25 * c2] = aa(c0, c1, c2) | (aa(comp(c2), comp(c1), comp(c0)) << 8)
    return lut


def fdiv(a, b):
    return a // b  # python floors; the kernels use an explicit floor division helper


def lines_stage(out, sym, lut, dbase, hdr_off, hdr_end, Lp, out_off, emitted, mask, alternative, tile, S2):
    """k_lines: block ranges per (tile, record piece, strand) with the kernel's tile-relative
    arithmetic (build step of gts_lines.cu), then every owned line written residue by residue."""
    nslots = len(dbase) - 1

    def symbol(e):
        return sym[e] if 0 <= e < len(sym) else 0  # the staged halo is zero past the end of the stream

    ntiles = (S2 + tile - 1) // tile
    for t in range(ntiles):
        T0 = t * tile
        nsym = min(T0 + tile, S2) - T0
        s = 0
        while dbase[s + 1] <= T0:
            s += 1
        while s < nslots and dbase[s] < T0 + nsym:
            D = dbase[s]
            n = dbase[s + 1] - 2 - D
            if emitted(s):
                H = hdr_end[s] - hdr_off[s]
                d = D - T0
                bodies, base = [], out_off[s]
                for f in range(6):
                    bodies.append(base + H + 3)
                    if (mask >> f) & 1:
                        base += run_size(H, Lp[s][f])
                for strand in (0, 1):
                    # the strand's frames by codon phase k
                    if strand == 0:
                        fr = [0, 1, 2]
                    else:
                        fr = [3 + (k if alternative else (n % 3 - k) % 3) for k in range(3)]
                    Lk = [Lp[s][f] if (mask >> f) & 1 else 0 for f in fr]
                    NLk = [(L + 59) // 60 for L in Lk]
                    NLmax = max(NLk)
                    lo = hi = 0
                    if NLmax:
                        if strand == 0:
                            x = -(d + 2)
                            if x > 0:
                                lo = (x - 1) // 180 + 1
                            e = nsym - (d + 2) - 180 * lo
                            hi = lo + ((e + 179) // 180 if e > 0 else 0)
                        else:
                            Lc = min((n - 189) // 180 + 1 if n >= 189 else 0, NLmax)
                            y = d + n - 189
                            if y >= 0:
                                qy = y // 180
                                ry = y - 180 * qy
                                hi = qy + 1
                                if nsym > ry:
                                    lo = max(qy + 1 - (nsym - ry + 179) // 180, 0)
                                else:
                                    lo = hi
                            hi = min(hi, Lc)
                            if Lc < NLmax and 0 <= d < nsym:
                                if hi <= lo:
                                    lo = Lc
                                hi = NLmax
                        hi = min(hi, NLmax)
                    for b in range(lo, hi):
                        for k in range(3):
                            if b >= NLk[k]:
                                continue
                            f = fr[k]
                            dst = bodies[f] + 61 * b
                            cnt = min(60, Lk[k] - 60 * b)
                            for x in range(cnt):
                                if strand == 0:
                                    q = k + 180 * b + 3 * x  # codon = nucleotides q, q+1, q+2
                                    idx = symbol(D + q) + 5 * symbol(D + q + 1) + 25 * symbol(D + q + 2)
                                    aa = lut[idx] & 0xFF
                                else:
                                    top = n - 1 - 180 * b - k - 3 * x  # read downwards: top, top-1, top-2
                                    idx = symbol(D + top - 2) + 5 * symbol(D + top - 1) + 25 * symbol(D + top)
                                    aa = lut[idx] >> 8
                                assert out[dst + x] == ord("?")
                                out[dst + x] = aa
                            assert out[dst + cnt] == ord("?")
                            out[dst + cnt] = 0x0A
            s += 1


def translate(raw, aa64, mask, clean=False, alternative=False, trim=False, tile=48, residues="windows"):
    assert tile % 3 == 0
    lut = build_lut(aa64, clean)
    sym, hdr_off, hdr_end, dbase, total = parse(raw)
    R = len(hdr_off) - 1
    nslots = R + 1
    S2 = total + 2

    # ---- k_size ----
    Lp, out_off = [], []
    off = 0
    for s in range(nslots):
        D = dbase[s]
        n = dbase[s + 1] - 2 - D
        emitted = s >= 1 or n > 0 or R == 0
        if trim:
            L = trimmed_lengths(sym, lut, D, n, alternative, mask)
        else:
            L = frame_lengths(n, alternative)
        Lp.append(L)
        out_off.append(off)
        if emitted:
            H = hdr_end[s] - hdr_off[s]
            off += sum(run_size(H, L[f]) for f in range(6) if (mask >> f) & 1)
    out = bytearray(b"?" * off)

    def emitted(s):
        n = dbase[s + 1] - 2 - dbase[s]
        return s >= 1 or n > 0 or R == 0

    # ---- residues ----
    ntiles = (S2 + tile - 1) // tile
    if residues == "lines":
        lines_stage(out, sym, lut, dbase, hdr_off, hdr_end, Lp, out_off, emitted, mask, alternative, tile, S2)
        ntiles = 0
    for t in range(ntiles):
        T0 = t * tile
        Tend = min(T0 + tile, S2)
        # phase arrays (windows ending at local symbol el = 3m + phi)
        F = [[0] * (tile // 3) for _ in range(3)]
        Rv = [[0] * (tile // 3) for _ in range(3)]
        for el in range(Tend - T0):
            e = T0 + el
            if e < 2:
                continue
            f_, r_ = window_aa(sym, lut, e)
            F[el % 3][el // 3] = f_
            Rv[el % 3][el // 3] = r_
        # first slot whose window-end range reaches into the tile
        s = 0
        while dbase[s + 1] <= T0:
            s += 1
        while s < nslots and dbase[s] < Tend:
            D = dbase[s]
            n = dbase[s + 1] - 2 - D
            if emitted(s):
                H = hdr_end[s] - hdr_off[s]
                q_lo = max(-2, T0 - D - 2)
                q_hi = min(n - 1, Tend - 1 - D - 2)
                run_base = out_off[s]
                for f in range(6):
                    if not (mask >> f) & 1:
                        continue
                    L = Lp[s][f]
                    body = run_base + H + 3
                    run_base += run_size(H, L)
                    if f < 3:
                        k = f
                        lo = max(q_lo, 0)
                        j_lo = max(0, -fdiv(-(lo - k), 3))  # ceil((lo-k)/3)
                        j_hi = fdiv(q_hi - k, 3) + 1 if q_hi >= k else 0
                        j_hi = min(j_hi, L)
                        if j_hi <= j_lo:
                            continue
                        el0 = D - T0 + k + 2 + 3 * j_lo
                        arr, m0, d = F[el0 % 3], el0 // 3, 1
                    else:
                        p = staden_start(n, f - 3, alternative)
                        Q = n - 3 - p
                        j_lo = max(0, -fdiv(-(Q - q_hi), 3))
                        j_hi = fdiv(Q - q_lo, 3) + 1 if Q >= q_lo else 0
                        j_hi = min(j_hi, L)
                        if j_hi <= j_lo:
                            continue
                        el0 = D - T0 + Q - 3 * j_lo + 2
                        arr, m0, d = Rv[el0 % 3], el0 // 3, -1
                    naa = j_hi - j_lo
                    col0 = j_lo % 60
                    g0 = body + j_lo + j_lo // 60
                    final = j_hi == L
                    nlines = (col0 + naa + 59) // 60
                    for ln in range(nlines):
                        start = 0 if ln == 0 else (60 - col0) + 60 * (ln - 1)
                        cnt = min(60 - (col0 if ln == 0 else 0), naa - start)
                        dst = g0 + start + ln
                        for i in range(cnt):
                            assert out[dst + i] == ord("?")
                            out[dst + i] = arr[m0 + d * (start + i)]
                        nl = ((col0 + start + cnt) % 60 == 0) or (final and start + cnt == naa)
                        if nl:
                            assert out[dst + cnt] == ord("?")
                            out[dst + cnt] = 0x0A
            s += 1

    # ---- k_headers ----
    for s in range(nslots):
        if not emitted(s):
            continue
        H = hdr_end[s] - hdr_off[s]
        hdr = raw[hdr_off[s]: hdr_off[s] + H]
        sp = hdr.find(b" ")
        if sp < 0:
            sp = H
        run_base = out_off[s]
        for f in range(6):
            if not (mask >> f) & 1:
                continue
            for i in range(H + 3):
                if i < sp:
                    c = hdr[i]
                elif i == sp:
                    c = ord("_")
                elif i == sp + 1:
                    c = ord("1") + f
                elif i < H + 2:
                    c = hdr[i - 2]
                else:
                    c = 0x0A
                assert out[run_base + i] == ord("?")
                out[run_base + i] = c
            run_base += run_size(H, Lp[s][f])
    return bytes(out)
