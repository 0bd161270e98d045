#!/usr/bin/env python3
"""Fixture generator (test infrastructure): random LUT/FF netlist in BLIF for the reference
flow.  Deterministic for a given seed (Python's Mersenne Twister).

usage: gen_blif.py OUT.blif --luts 300 --pis 16 --window 60 --seed 1
"""
import argparse
import random


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--luts", type=int, default=300)
    ap.add_argument("--pis", type=int, default=16)
    ap.add_argument("--window", type=int, default=60)
    ap.add_argument("--latch_frac", type=float, default=0.3)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--name", default="toy")
    ap.add_argument("--hub", type=int, default=0, help="number of LUTs that additionally read one shared hub signal (a high-fanout net)")
    ap.add_argument("--clocks", type=int, default=1, help="number of clock domains the latches are spread over")
    ap.add_argument("--mults", type=int, default=0, help="number of `.subckt mult4` instances (4x4 -> 8 bit hard blocks of the heterogeneous "
                    "fixture architecture k6_N10_het.xml), spread evenly through the LUT sequence; 0 keeps the output of older versions")
    a = ap.parse_args()
    rng = random.Random(a.seed)
    signals = ["pi%d" % i for i in range(a.pis)]
    used = set()
    body = []
    latches = 0
    mult_at = {(j + 1) * a.luts // (a.mults + 1): j for j in range(a.mults)}
    for i in range(a.luts):
        if i in mult_at and len(signals) >= 8:
            j = mult_at[i]
            ins = rng.sample(signals[-a.window:], 8)
            used.update(ins)
            pins = ["a[%d]=%s" % (b, ins[b]) for b in range(4)] + ["b[%d]=%s" % (b, ins[4 + b]) for b in range(4)]
            pins += ["out[%d]=m%d_%d" % (b, j, b) for b in range(8)]
            body.append(".subckt mult4 " + " ".join(pins))
            signals.extend("m%d_%d" % (j, b) for b in range(8))
        k = rng.randint(3, 6)
        pool = signals[-a.window:]
        ins = rng.sample(pool, min(k, len(pool)))
        if i < a.hub and "pi0" not in ins:
            ins = ins[:5] + ["pi0"]          # pi0 becomes a net with >= 64 sinks (HIGH_FANOUT_NET_LIM, vpr_types.h:91)
        used.update(ins)
        out = "n%d" % i
        body.append(".names %s %s" % (" ".join(ins), out))
        # random single-cube cover with at least one care literal
        cube = "".join(rng.choice("01-") for _ in ins)
        if set(cube) == {"-"}:
            cube = "1" + cube[1:]
        body.append("%s 1" % cube)
        if rng.random() < a.latch_frac:
            q = "q%d" % i
            body.append(".latch %s %s re %s 0" % (out, q, "clk" if a.clocks <= 1 else "clk%d" % (latches % a.clocks)))
            used.add(out)
            signals.append(q)
            latches += 1
        else:
            signals.append(out)
    outs = [s for s in signals if s not in used and not s.startswith("pi")]
    # unused primary inputs would be dangling: feed each into an extra output buffer-less PO is illegal,
    # so only declare the primary inputs that are actually used.
    pis = [s for s in signals[:a.pis] if s in used]
    with open(a.out, "w") as f:
        f.write(".model %s\n" % a.name)
        clk_names = (["clk"] if a.clocks <= 1 else ["clk%d" % c for c in range(min(a.clocks, latches))]) if latches else []
        f.write(".inputs %s\n" % " ".join(pis + clk_names))
        f.write(".outputs %s\n" % " ".join(outs))
        f.write("\n".join(body))
        f.write("\n.end\n")
        if a.mults:
            f.write("\n.model mult4\n.inputs %s\n.outputs %s\n.blackbox\n.end\n" % (
                " ".join(["a[%d]" % b for b in range(4)] + ["b[%d]" % b for b in range(4)]), " ".join("out[%d]" % b for b in range(8))))
    print("wrote %s: %d luts, %d latches, %d inputs, %d outputs" % (a.out, a.luts, latches, len(pis), len(outs)))


if __name__ == "__main__":
    main()
