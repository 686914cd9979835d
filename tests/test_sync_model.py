"""CPU: a bit-level model of the decode side's self-synchronisation (csrc/tic_decode.cu, phases 1-3), run on
streams of the CPU restatement.  It pins two properties of the algorithm that do not depend on the GPU:

  * the fixed point of `E[g+1] = decode(g, E[g]).exit` with a known E[first] is the serial parse — block counts
    and DC-difference sums per subsequence add up to the coefficients encode() produced;
  * the two rules that keep decode(entry) a function on garbage parses (no codeword -> skip one bit; more than 63
    coefficients without an EOB -> restart one bit further on, expecting a DC symbol) make guessed entries fall
    into step quickly even on periodic bit patterns (flat images), where a wrong phase can otherwise parse the
    stream as an endless run of AC symbols.

The model mirrors decode_sub(): same state (bit overshoot, zigzag index), same rules, same 1024-bit subsequences."""
import numpy as np
import pytest

from oracle import oracle_lib as O
from tests.cases import make_case, synthetic_image

SUB = 1024


def _tables():
    dc = {O.default_code(False, s): s for s in range(12)}
    ac = {}
    for s in range(256):
        c = O.default_code(True, s)
        if c is not None:
            ac[c] = s
    return dc, ac


DC_TAB, AC_TAB = _tables()


def _lookup(bits, p, table):
    """read_huffman_code (huffman.py:66-74): the shortest matching prefix of at most 16 bits, or None."""
    for ln in range(1, 17):
        s = table.get(bits[p:p + ln]) if p + ln <= len(bits) + 16 else None
        if s is not None and len(bits[p:p + ln]) == ln:
            return s, ln
    return None


def decode_sub(bits, start, entry, restart_rule=True):
    """(exit state, blocks started, DC-difference sum) of one subsequence; bits past the end read as zeros."""
    p, z = start + entry[0], entry[1]
    end = min(start + SUB, len(bits))
    padded = bits + "0" * 64
    n = dsum = 0
    it = 0
    while p < end and it < 4096:
        it += 1
        hit = _lookup(padded, p, DC_TAB if z == 0 else AC_TAB)
        if hit is None:
            p += 1
            continue
        sym, ln = hit
        size = sym & 15
        val = 0
        if size:
            vb = int(padded[p + ln:p + ln + size], 2)
            val = vb if vb >> (size - 1) else vb - (1 << size) + 1
        p0 = p
        p += ln + size
        if z == 0:
            n += 1
            dsum += val
            z = 1
        elif sym == 0:
            z = 0
        else:
            z += sym >> 4
            if z > 63:
                if restart_rule:
                    z, p = 0, p0 + 1
                    continue
                z = 200   # without the rule: keep going as the reference would, never storing anything
            else:
                z += 1
    return (max(p - (start + SUB), 0), z), n, dsum


def synchronise(bits, restart_rule=True, max_rounds=10_000):
    """Jacobi rounds over the subsequences of one stream (data starts at bit 128); returns rounds, counts, sums."""
    nsubs = (len(bits) - 128 + SUB - 1) // SUB
    E = [(0, 0)] * nsubs
    used = [None] * nsubs
    nd = [(0, 0)] * nsubs
    rounds = 0
    while True:
        rounds += 1
        changed = False
        newE = list(E)
        for g in range(nsubs):
            if used[g] == E[g]:
                continue
            used[g] = E[g]
            ex, n, d = decode_sub(bits, 128 + g * SUB, E[g], restart_rule)
            nd[g] = (n, d)
            if g + 1 < nsubs and newE[g + 1] != ex:
                newE[g + 1] = ex
                changed = True
        E = newE
        if not changed or rounds >= max_rounds:
            return rounds, nd


def _bits(stream):
    return "".join(format(b, "08b") for b in stream)


@pytest.mark.parametrize("kind,q", [("synthetic", 50), ("synthetic", 90), ("noise", 50), ("impulse", 75), ("binary", 20)])
def test_fixed_point_is_the_serial_parse(kind, q):
    img = synthetic_image(96, 160, 3) if kind == "synthetic" else make_case({"kind": kind, "shape": (96, 160), "seed": 9})
    stream = O.compress(img, q)
    rounds, nd = synchronise(_bits(stream))
    e = O.encode(img, q)
    nblk = len(e["dc"])
    # blocks started per subsequence: only padding can add phantom blocks, and only in the last subsequence
    assert sum(n for n, _ in nd[:-1]) <= nblk <= sum(n for n, _ in nd)
    # running DC sums at every subsequence boundary agree with np.cumsum of encode()'s differences
    dc_cum = np.concatenate([[0], np.cumsum(e["dc"])])
    blocks = dsum = 0
    for n, d in nd[:-1]:
        blocks += n
        dsum += d
        assert dsum == dc_cum[blocks]
    assert rounds <= 6


@pytest.mark.parametrize("value", list(range(0, 256, 15)) + [77, 128, 255])
def test_flat_images_synchronise_in_a_few_rounds(value):
    """A flat image is one short block repeated; from some phases a guessed entry parses it as AC symbols for ever.
    With the restart rule every guess falls into step inside its own subsequence."""
    stream = O.compress(np.full((256, 256), value, np.uint8), 50)
    rounds, nd = synchronise(_bits(stream))
    assert rounds <= 4, rounds
    assert sum(n for n, _ in nd) >= 1024


def test_without_the_restart_rule_a_flat_image_is_serial():
    """The reason for the rule: without it the correction has to travel through the stream one subsequence per
    round."""
    stream = O.compress(np.full((512, 512), 77, np.uint8), 50)
    bits = _bits(stream)
    nsubs = (len(bits) - 128 + SUB - 1) // SUB
    with_rule, _ = synchronise(bits, restart_rule=True)
    without, _ = synchronise(bits, restart_rule=False)
    assert with_rule <= 4
    assert without >= nsubs // 2, (without, nsubs)


# ---- the planned early stop of a repeated decode ("What comes next") ---------------------------------
# A decode from a corrected entry meets the parse of the previous decode after a few dozen bits.  If the previous
# decode left a mask of its DC-symbol starts, the repeat can stop at the first DC start the two parses share and
# take the rest of its result from the previous one.  The model below checks the bookkeeping of that plan (the
# mask update, the block count, the DC sum) against full decodes; the CUDA side does not use it yet.

def _dc_value(padded, p):
    sym, ln = _lookup(padded, p, DC_TAB)
    size = sym & 15
    if not size:
        return 0
    vb = int(padded[p + ln:p + ln + size], 2)
    return vb if vb >> (size - 1) else vb - (1 << size) + 1


def decode_sub_early(bits, start, entry, prev):
    """prev: None or dict(mask=set of DC-start positions, n, dsum, exit) of the previous decode of this subsequence.
    Returns the same dict for this decode, and the number of bits actually parsed."""
    p, z = start + entry[0], entry[1]
    end = min(start + SUB, len(bits))
    padded = bits + "0" * 64
    mask = set(prev["mask"]) if prev else set()
    n = dsum = removed_n = removed_d = 0
    first = p

    def drop(lo, hi):   # DC starts of the previous parse that the new parse passes over
        nonlocal removed_n, removed_d
        for b in [b for b in mask if lo <= b < hi]:
            mask.discard(b)
            removed_n += 1
            removed_d += _dc_value(padded, b)

    drop(start, p)
    it = 0
    while p < end and it < 4096:
        it += 1
        if z == 0 and prev and p in mask:   # the two parses are in the same state: the rest is identical
            return dict(mask=mask, n=n + prev["n"] - removed_n, dsum=dsum + prev["dsum"] - removed_d,
                        exit=prev["exit"]), p - first
        hit = _lookup(padded, p, DC_TAB if z == 0 else AC_TAB)
        if hit is None:
            drop(p, p + 1)
            p += 1
            continue
        sym, ln = hit
        size = sym & 15
        val = 0
        if size:
            vb = int(padded[p + ln:p + ln + size], 2)
            val = vb if vb >> (size - 1) else vb - (1 << size) + 1
        p0 = p
        p += ln + size
        if z == 0:
            drop(p0 + 1, p)
            mask.add(p0)
            n += 1
            dsum += val
            z = 1
        elif sym == 0:
            drop(p0, p)
            z = 0
        else:
            z += sym >> 4
            if z > 63:
                z, p = 0, p0 + 1
                drop(p0, p)
                continue
            drop(p0, p)
            z += 1
    drop(min(p, end), end)
    return dict(mask=mask, n=n, dsum=dsum, exit=(max(p - (start + SUB), 0), z)), min(p, end) - first


@pytest.mark.parametrize("kind,q", [("synthetic", 50), ("synthetic", 92), ("noise", 50), ("flat", 50), ("impulse", 30)])
def test_early_stop_bookkeeping_matches_full_decodes(kind, q):
    img = synthetic_image(128, 192, 7) if kind == "synthetic" else \
        make_case({"kind": kind, "shape": (128, 192), "seed": 3, "value": 77})
    bits = _bits(O.compress(img, q))
    nsubs = (len(bits) - 128 + SUB - 1) // SUB
    rng = np.random.default_rng(q)
    parsed = full = 0
    for g in range(1, nsubs):
        start = 128 + g * SUB
        state = None
        # a chain of repeated decodes from different entries, the way repair rounds produce them: the guess first,
        # then a few arbitrary (wrong) entries, then anything again
        entries = [(0, 0)] + [(int(rng.integers(0, 31)), int(rng.integers(0, 20))) for _ in range(3)]
        for e in entries:
            state, nbits = decode_sub_early(bits, start, e, state)
            ex, n, d = decode_sub(bits, start, e)
            assert (state["exit"], state["n"], state["dsum"]) == (ex, n, d), (g, e)
            if e != (0, 0):
                parsed += nbits
                full += min(SUB, len(bits) - start) - e[0]
    assert parsed < 0.6 * full   # repeats stop early (on these small images; ~4-15 % on the benchmark's streams)
