"""Host model of the per-lane stream ring of the one-thread decoders (zipnn_b200/csrc/decode.cuh: ring_top_up,
window_refill, decode16).  The device code waits with cp.async.wait_group 2, i.e. a block requested in half-iteration
g is only guaranteed to have landed at the END of half g + 2: this checks that no refill ever reads a ring word before
that, for random and adversarial code lengths, and that a slot is never overwritten while its bytes are still unread."""
import random

RING = 64


def simulate(lengths, start_consumed, seed_span):
    # offsets grow upwards; the stream is read downwards.  qm = offset of the word the next refill reads.
    qm = 10_000 * 4
    consumed = start_consumed          # bits consumed from the 64-bit container (after a refill: < 32)
    fetch = qm + 4 - seed_span         # lowest requested byte: span = qm + 4 - fetch
    assert fetch % 16 == 0
    landed_by_half = {}                # block start -> half in which it was requested (seed blocks: long ago)
    for f in range(fetch, qm + 64, 16):
        landed_by_half[f] = -10
    half = 0
    it = iter(lengths)
    try:
        while True:
            # ring_top_up(b, 1)
            f = fetch - 16
            if f + (RING - 4) >= qm:
                # the slot being overwritten holds [f + 64, f + 80): all of it must be dead (>= qm + 4)
                assert f + 64 >= qm + 4
                landed_by_half[f] = half
                fetch = f
            for _pair in range(4):     # 8 symbols, one refill test per pair (peek-before-refill does not change qm)
                if consumed >= 32:
                    consumed -= 32
                    word = qm          # the refill reads [qm, qm + 4)
                    blk = word & ~15
                    assert blk in landed_by_half, "read below the fetch frontier"
                    assert landed_by_half[blk] <= half - 3, (half, landed_by_half[blk])
                    qm -= 4
                for _ in range(2):
                    consumed += next(it)
                    assert consumed <= 53
            half += 1                  # cp.async.wait_group 2: groups <= half - 3 have landed for the next half
    except StopIteration:
        return half


def test_ring_lead_adversarial():
    for start in range(0, 32):
        for span in range(40, 64, 4):
            if (40_000 + 4 - span) % 16:
                continue
            assert simulate([11] * 40_000, start, span) > 1000
            assert simulate([1] * 40_000, start, span) > 1000


def test_ring_lead_random():
    rng = random.Random(7)
    for _ in range(200):
        n = 20_000
        mode = rng.random()
        if mode < 0.3:
            lens = [rng.choice((10, 11)) for _ in range(n)]
        elif mode < 0.6:
            lens = [rng.randint(1, 11) for _ in range(n)]
        else:  # bursts of long codes between short ones
            lens = []
            while len(lens) < n:
                lens += [rng.randint(1, 3)] * rng.randint(1, 40) + [11] * rng.randint(1, 60)
        span = rng.choice([s for s in range(40, 64, 4) if (40_000 + 4 - s) % 16 == 0])
        simulate(lens[:n], rng.randint(0, 31), span)
