"""Single-warp event-step model of a SASS loop (B300_MICROARCH.md 'Single-warp T_1w'): walks the instructions of a
loop body in program order for several iterations, honouring stall counts, scoreboard waits and a simple XU-pipe queue,
and reports cycles per iteration plus where the exposed scoreboard waits are.

    cuobjdump -sass -fun <mangled> lib.so > k.sass
    python tools/sass_sim.py k.sass 0x1b50 0x5020 [--packed2] [--skip lo:hi ...]

Branches inside the range are assumed NOT taken except the loop-closing one (the rare-event blocks of the fill kernels
are skipped with --skip, giving their address ranges).
"""
import re, sys, argparse, collections

LAT = {"MUFU": 18, "LDS": 29, "SHFL": 24, "LDG": 300, "LDC": 30, "STG": 10, "LDGSTS": 300, "F2F": 12, "DEPBAR": 0}
XU_RT = 8          # cycles of XU pipe per warp-wide MUFU (measured: 0.516 warp-instr/clk/SM)


def parse(path):
    lines = open(path).read().split("\n")
    pat = re.compile(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/")
    pat2 = re.compile(r"/\* 0x([0-9a-f]{16}) \*/")
    ins, i = [], 0
    while i < len(lines):
        m = pat.search(lines[i])
        if m and i + 1 < len(lines):
            m2 = pat2.search(lines[i + 1])
            if m2:
                ctrl = (int(m2.group(1), 16) >> 41) & 0x1FFFFF
                text = m.group(2).strip()
                op = (text.split()[1] if text.startswith("@") else text.split()[0]).split(".")[0]
                ins.append(dict(addr=int(m.group(1), 16), text=text, op=op, stall=ctrl & 0xF, wb=(ctrl >> 5) & 7,
                                rb=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 0x3F))
                i += 2
                continue
        i += 1
    return ins


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sass"); ap.add_argument("lo"); ap.add_argument("hi")
    ap.add_argument("--packed2", action="store_true", help="packed f32x2 ops block issue for 2 cycles")
    ap.add_argument("--skip", nargs="*", default=[], help="address ranges lo:hi (hex) to skip (rare blocks)")
    ap.add_argument("--iters", type=int, default=6)
    ap.add_argument("--top", type=int, default=12)
    a = ap.parse_args()
    lo, hi = int(a.lo, 16), int(a.hi, 16)
    skips = [tuple(int(x, 16) for x in s.split(":")) for s in a.skip]
    body = [x for x in parse(a.sass) if lo <= x["addr"] <= hi and not any(s0 <= x["addr"] < s1 for s0, s1 in skips)]
    T = 0.0
    sb = [0.0] * 6
    xu_free = 0.0
    per_iter = []
    exposed = collections.Counter()
    prev_stall = 0
    for it in range(a.iters):
        t_start = T
        for x in body:
            t_ready = T + prev_stall
            t_arm = max([sb[s] for s in range(6) if (x["wait"] >> s) & 1], default=0.0)
            if t_arm > t_ready and it == a.iters - 1:
                exposed[(x["addr"], x["text"][:60])] += t_arm - t_ready
            T = max(t_ready, t_arm)
            st = x["stall"]
            if a.packed2 and x["op"] in ("FFMA2", "FMUL2", "FADD2"):
                st = max(st, 2)
            prev_stall = st
            lat = LAT.get(x["op"], None)
            if x["op"] == "MUFU":
                start = max(T, xu_free)
                xu_free = start + XU_RT
                done = start + LAT["MUFU"]
            else:
                done = T + (lat if lat is not None else 6)
            if x["wb"] < 6:
                sb[x["wb"]] = max(sb[x["wb"]], done)
            if x["rb"] < 6:
                sb[x["rb"]] = max(sb[x["rb"]], T + 6)
        per_iter.append(T - t_start)
    n = len(body)
    print(f"{n} instructions per iteration; cycles per iteration: {[round(v) for v in per_iter]}")
    print("exposed scoreboard waits in the last iteration (cycles):")
    for (addr, text), c in exposed.most_common(a.top):
        print(f"  {addr:05x}  {c:6.0f}  {text}")
    print(f"  total exposed {sum(exposed.values()):.0f}")


if __name__ == "__main__":
    main()
