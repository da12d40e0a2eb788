"""Summarise an `ncu --page source --csv --print-source=sass` export: opcode mix and stall reasons.

usage: python tools/ncu_sass_summary.py <file.csv> [top]
"""
import io
import sys

import pandas as pd


def main(path, top=18):
    lines = open(path).read().split("\n")
    df = pd.read_csv(io.StringIO("\n".join(lines[1:])), dtype=str)
    num = lambda c: pd.to_numeric(df[c], errors="coerce").fillna(0)  # noqa: E731
    df["inst"] = num("Instructions Executed")
    df["samples"] = num("# Samples")
    src = df["Source"].str.strip().str.replace(r"^@!?U?P\d+\s+", "", regex=True)
    df["op"] = src.str.split().str[0].str.split(".").str[0]
    total = df["inst"].sum()
    print(f"SASS lines {len(df)}  warp instructions {total:.0f}  samples {df['samples'].sum():.0f}")
    g = df.groupby("op").agg(inst=("inst", "sum"), samples=("samples", "sum")).sort_values("inst", ascending=False)
    g["pct_inst"] = (100 * g.inst / total).round(1)
    g["pct_samples"] = (100 * g.samples / df["samples"].sum()).round(1)
    print(g.head(top).to_string())
    # Blackwell / Hopper data-movement and synchronisation opcodes (static SASS lines | executed warp instructions)
    full = src.str.split().str[0]
    for label, pattern in (("UBLKCP (TMA bulk copy)", "UBLKCP"), ("LDGSTS (cp.async)", "LDGSTS"), ("SYNCS (mbarrier)", "SYNCS"),
                           ("ATOMG / RED (global atomics)", "ATOMG|RED"), ("ATOMS (shared atomics)", "ATOMS"), ("BAR (block barrier)", "BAR"),
                           ("LDS / STS", "LDS|STS"), ("LDG / STG", "LDG|STG"), ("MUFU", "MUFU"), ("VOTE / REDUX", "VOTE|REDUX|CREDUX")):
        mask = full.str.contains("^(?:" + pattern + ")", regex=True, na=False)
        print(f"  {label:34s} {int(mask.sum()):5d} lines  {df.loc[mask, 'inst'].sum():14.0f} executed")
    stall_cols = [c for c in df.columns if c.startswith("stall_") and "Not Issued" not in c]
    stalls = pd.Series({c: num(c).sum() for c in stall_cols}).sort_values(ascending=False)
    print((100 * stalls / stalls.sum()).round(1).head(8).to_string())


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 18)


def segments(path):
    """Instruction / sample totals between consecutive BAR.SYNC / SYNCS (phase boundaries), in address order."""
    lines = open(path).read().split("\n")
    df = pd.read_csv(io.StringIO("\n".join(lines[1:])), dtype=str)
    inst = pd.to_numeric(df["Instructions Executed"], errors="coerce").fillna(0)
    samp = pd.to_numeric(df["# Samples"], errors="coerce").fillna(0)
    src = df["Source"].str.strip()
    seg, acc_i, acc_s, n, fp = 0, 0.0, 0.0, 0, 0.0
    for s, i, k in zip(src, inst, samp):
        acc_i += i
        acc_s += k
        n += 1
        if any(op in s for op in ("DFMA", "DMUL", "DADD")):
            fp += i
        if "BAR.SYNC" in s or "SYNCS.PHASECHK" in s:
            print(f"segment {seg}: {n:5d} SASS lines  {acc_i/1e6:8.2f} M warp-instr ({100*fp/max(acc_i,1):4.1f}% fp64)  {acc_s:7.0f} samples   ends at: {s[:50]}")
            seg, acc_i, acc_s, n, fp = seg + 1, 0.0, 0.0, 0, 0.0
    print(f"segment {seg}: {n:5d} SASS lines  {acc_i/1e6:8.2f} M warp-instr ({100*fp/max(acc_i,1):4.1f}% fp64)  {acc_s:7.0f} samples   (tail)")
