"""Host model of the re-scan shortcut of the per-bitstream-CTA decoder (zipnn_b200/csrc/decode_sync.cuh: swin_scan records
every eighth code boundary of a thread's first pass over its segment, swin_rescan decodes from a corrected start only
until it stands on a recorded boundary).  Decoding is deterministic, so from a shared boundary both passes coincide:
the shortcut must return exactly what a full pass from the corrected start returns (stop position and symbol count)."""
import random

CHECK_EVERY = 8
MAX_CP = 32


def make_code(rng, nsym):
    """A complete prefix code: lengths by splitting leaves at random; -> decode table {bitstring: symbol}."""
    leaves = [""]
    while len(leaves) < nsym:
        i = rng.randrange(len(leaves))
        if len(leaves[i]) >= 11:
            continue
        s = leaves.pop(i)
        leaves += [s + "0", s + "1"]
    return {code: k for k, code in enumerate(leaves)}


def next_len(bits, pos, table):
    """Length of the code that starts at bit `pos` (bits are consumed towards higher indices here)."""
    for ln in range(1, 12):
        if bits[pos:pos + ln] in table:
            return ln
    raise AssertionError("complete code: cannot happen inside the padded stream")


def scan(bits, table, start, bound):
    """First pass: from `start` to the first boundary at or past `bound`.  -> (stop, n, checkpoints)."""
    pos, n, cps = start, 0, []
    while pos < bound:
        pos += next_len(bits, pos, table)
        n += 1
        if n % CHECK_EVERY == 0 and pos < bound and len(cps) < MAX_CP:
            cps.append(pos)
    return pos, n, cps


def rescan(bits, table, start, bound, cps, n_rec, stop_rec):
    """From another start: stop at the first recorded boundary (-> recorded stop, adjusted count) or run to the bound."""
    pos, n, j = start, 0, 0
    while pos < bound:
        pos += next_len(bits, pos, table)
        n += 1
        if pos >= bound:
            break
        while j < len(cps) and cps[j] < pos:
            j += 1
        if j < len(cps) and cps[j] == pos:
            return stop_rec, n + n_rec - CHECK_EVERY * (j + 1), True
    return pos, n, False


def test_rescan_equals_full_scan():
    rng = random.Random(11)
    shortcuts = 0
    for trial in range(300):
        table = make_code(rng, rng.choice([2, 3, 8, 40, 200]))
        codes = list(table)
        weights = [2.0 ** -len(c) for c in codes]
        nsym = rng.randint(50, 3000)
        bits = "".join(rng.choices(codes, weights, k=nsym)) + "0" * 16   # padding: an overshooting pass stays readable
        total = len(bits) - 16
        lo = rng.randint(0, max(0, total - 200))
        bound = min(total, lo + rng.randint(30, 900))
        guess = lo                                      # a thread's first start: a guess (any bit position)
        stop1, n1, cps = scan(bits, table, guess, bound)
        for _ in range(4):                              # corrected starts at or before the guess, as in the kernel's rounds
            start2 = rng.randint(max(0, lo - 40), lo)
            want_stop, want_n, _ = scan(bits, table, start2, bound)
            got_stop, got_n, hit = rescan(bits, table, start2, bound, cps, n1, stop1)
            assert (got_stop, got_n) == (want_stop, want_n), (trial, start2, hit)
            shortcuts += hit
    assert shortcuts > 100   # the shortcut is actually taken (codes synchronise)
