"""The algebra behind envelope_kernel (mixlab_b200/csrc/envelope.cu), checked on the CPU against the reference's state
machine (src/module/envelope.rs:96-117): a run of gate samples is summarised by Seg = {first event, last event, latest two
transitions among the events after the first}; runs are joined with seg_combine; the machine state before any sample follows
from the Seg of everything before it and the class the call started in.  The Python below restates the kernel's
seg_combine / resolution line by line -- it is the host-side mirror of the device logic, not the oracle."""
import numpy as np
import pytest


def key_of(idx, on):
    return ((idx + 1) << 1) | (1 if on else 0)


def seg_of(gate, lo, hi):
    """Seg of samples [lo, hi): (F, L, a, b), keys as in the kernel, 0 = none."""
    F = L = a = b = 0
    prev_cls = None
    for i in range(lo, hi):
        x = gate[i]
        if x == 1.0 or x == 0.0:
            on = x == 1.0
            k = key_of(i, on)
            if F == 0:
                F = k
            elif on != prev_cls:                       # an event after the first whose class differs from the event before
                b, a = a, k
            L = k
            prev_cls = on
    return (F, L, a, b)


def seg_combine(x, y):
    """envelope.cu seg_combine: x earlier in time than y."""
    xF, xL, xa, xb = x
    yF, yL, ya, yb = y
    xh, yh = xL != 0, yL != 0
    boundary = yF if (xh and ((xL ^ yF) & 1)) else 0
    F = xF if xh else yF
    L = yL if yh else xL
    third = boundary if boundary else xa
    fourth = xa if boundary else xb
    a = ya if ya else third
    b = (yb if yb else third) if ya else fourth
    return (F, L, a, b)


def resolve(P, c0):
    """Class just before the next sample and the latest two transitions before it, given the Seg of everything before it
    in the call and the class the call starts in (env_emit)."""
    F, L, a, b = P
    cls = bool(L & 1) if L else c0
    first_tr = F if (F and bool(F & 1) != c0) else 0
    ta = a if a else first_tr
    tb = (b if b else first_tr) if a else 0
    return cls, ta, tb


def machine(gate, c0):
    """The reference's transitions, sample by sample: yields (class before sample i, latest transition key, the one before)."""
    cls, ta, tb = c0, 0, 0
    out = []
    for i, x in enumerate(gate):
        out.append((cls, ta, tb))
        if not cls and x == 1.0:
            cls, tb, ta = True, ta, key_of(i, True)
        elif cls and x == 0.0:
            cls, tb, ta = False, ta, key_of(i, False)
    return out


def random_gate(rng, n, density):
    g = rng.uniform(0.01, 0.99, n).astype(np.float32)
    pos = rng.integers(0, n, max(1, int(n * density)))
    g[pos] = rng.integers(0, 2, pos.size).astype(np.float32)
    g[rng.integers(0, n, max(1, pos.size // 8))] = -0.0       # -0.0 == 0.0 is an off event
    return g


@pytest.mark.parametrize("density", [0.002, 0.05, 0.6])
def test_joining_pieces_gives_the_seg_of_the_whole(density):
    rng = np.random.default_rng(int(density * 1000))
    for _ in range(30):
        n = int(rng.integers(1, 400))
        g = random_gate(rng, n, density)
        cuts = sorted(set(rng.integers(0, n + 1, int(rng.integers(0, 12))).tolist()) | {0, n})
        acc = (0, 0, 0, 0)
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            acc = seg_combine(acc, seg_of(g, lo, hi))
        assert acc == seg_of(g, 0, n)
        # and in any bracketing: fold from the right
        acc = (0, 0, 0, 0)
        for lo, hi in reversed(list(zip(cuts[:-1], cuts[1:]))):
            acc = seg_combine(seg_of(g, lo, hi), acc)
        assert acc == seg_of(g, 0, n)


@pytest.mark.parametrize("c0", [False, True])
def test_prefix_seg_resolves_to_the_machine_state(c0):
    rng = np.random.default_rng(7 + c0)
    for density in (0.003, 0.04, 0.5):
        g = random_gate(rng, 600, density)
        want = machine(g, c0)
        for i in range(0, 600, 7):
            assert resolve(seg_of(g, 0, i), c0) == want[i], (density, i)


def test_a_saturated_seg_ignores_everything_earlier():
    """Two transitions inside a Seg: joined behind anything, its last event and its two transitions stand -- which is what
    lets the look-back stop there, and lets a saturated tile publish its own summary as inclusive."""
    rng = np.random.default_rng(3)
    g = random_gate(rng, 300, 0.3)
    for lo in range(0, 200, 13):
        y = seg_of(g, lo, lo + 90)
        if y[3] == 0:
            continue
        for xlo in range(0, lo, 17):
            r = seg_combine(seg_of(g, xlo, lo), y)
            assert r[1:] == y[1:]
            # and resolution never reads F of a saturated Seg
            assert resolve(r, False)[1:] == resolve(r, True)[1:] == (y[2], y[3])
