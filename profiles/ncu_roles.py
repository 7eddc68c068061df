#!/usr/bin/env python
"""Warp-stall samples of `lrf::render_kernel_t` by WARP ROLE (producers / MMA issuer / consumers) from an
.ncu-rep captured with `--set full --import-source on`:

    python profiles/ncu_roles.py gpurun_out/r2_fwd_cfg2_bf16.ncu-rep

The kernel is warp-specialised, so one SASS instruction belongs to one role.  Roles are told apart by landmarks in
the SASS listing (`ncu --page source --print-source sass --csv`), in program order:
  producers  from the CTA-wide prologue barrier to the first UTCHMMA-carrying block (the issuer's loop)
  issuer     the block holding the UTCHMMA instructions, up to the consumers' polling loop on `mma1`
  consumers  from that polling loop / the consumers' named barrier (BAR.SYNC 0x2) to the final BAR.SYNC 0x0
  tail       the final BAR.SYNC (every warp waits there for the slowest one)
and inside a role the waits are recognised by their instruction: NANOSLEEP loops (producers waiting for a free
A-tile buffer), SYNCS...TRYWAIT + branch (mbarrier waits).  Used for profiles/r2_forward_bf16_ncu.md."""
import csv
import io
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    ins = []
    for r in rows:
        if len(r) > 5 and r[0].startswith("0x") and r[4].isdigit():
            ins.append((len(ins), r[1].strip(), int(r[4]), int(r[5]) if r[5].isdigit() else 0))
    return ins


def main():
    ins = load(sys.argv[1])
    total = sum(x[2] for x in ins)
    idx = lambda key: [i for i, s, _, _ in ins if key in s]
    mma = idx("UTCHMMA")
    named = [i for i in idx("BAR.SYNC") if "0x2" in ins[i][1]]
    final = [i for i in idx("BAR.SYNC") if i > mma[-1] and "0x2" not in ins[i][1]]
    if not mma or not named or not final:
        sys.exit("landmarks not found (not a render_kernel_t capture?)")
    cons_bar, tail_bar = named[0], final[0]
    # the consumers' polling loop on mma1 sits right before their named barrier: walk back to the previous UTCBAR/commit
    commits = [i for i in idx("UTCBAR") if i < cons_bar]
    cons_lo = commits[-1] + 1
    issuer_lo = max(i for i in idx("SYNCS") if i < mma[0]) - 10
    tail_hi = tail_bar + 3
    roles = {"producers": (0, issuer_lo), "issuer": (issuer_lo, cons_lo), "consumers": (cons_lo, tail_bar - 1),
             "tail (final __syncthreads, all warps)": (tail_bar - 1, tail_hi), "outlined helpers": (tail_hi, len(ins))}
    print(f"total samples {total}, {len(ins)} SASS instructions")
    for name, (lo, hi) in roles.items():
        seg = ins[lo:hi]
        s = sum(x[2] for x in seg)
        print(f"\n{name}: instructions [{lo},{hi}) samples {s} = {100.0 * s / total:.1f} % of all, "
              f"{sum(x[3] for x in seg) / 1e6:.1f} M warp instructions executed")
        spin = 0
        for i, text, sm, ex in seg:
            if "NANOSLEEP" in text and ex > 0:
                spin += sum(x[2] for x in ins[i - 6:i + 6])
        if spin:
            print(f"    back-off spin loops (waiting for a free A-tile buffer): {spin} samples = {100.0 * spin / max(s, 1):.0f} % of the role")
        for i, text, sm, ex in sorted(seg, key=lambda x: -x[2])[:6]:
            print(f"    #{i:<6d} {sm:6d} samples  exec {ex:9d}  {text[:80]}")


if __name__ == "__main__":
    main()
