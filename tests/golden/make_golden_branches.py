#!/usr/bin/env python
"""Golden vectors for the two order-dependent steps `--mode build` runs after the load (run_laSV.sh:178):
the cleaning (`--remove_low_coverage_supernodes T`) and branch printing (./branches.fa).  Runs the compiled,
unmodified reference (oracle/_ref/dB_graph_*, built by oracle/Makefile from /root/reference) on the inputs of
the golden cases and stores, per (case, T): size, record count and SHA-256 of branches.fa, and record count and
SHA-256 of the records of the `.ctx` the same run dumps (in table order).  The files themselves are a few
hundred KB each and fully determined by the committed inputs, so only their digests are kept.

    python tests/golden/make_golden_branches.py        # needs /root/reference (this container only)
"""
import hashlib
import json
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from oracle import oracle as O  # noqa: E402
from tests import util  # noqa: E402

CASES = [("synth_cfg1_k31", None), ("synth_cfg0_k53", None), ("synth_cfg2_k95", None), ("synth_full_k53", None),
         ("synth_cfg1_k31", 1), ("synth_cfg0_k53", 1), ("synth_cfg2_k95", 1), ("synth_cfg0_k53", 2),
         ("edge_pe_k21_q5", None), ("edge_pe_k31_q5", None), ("edge_pe_k53_q5", None), ("edge_pe_k95_q5", None),
         ("edge_pe_k21_q5", 1)]


def inputs(name, wd):
    meta = util.case_meta(name)
    if "synth" in meta:
        seq, qual, _ = util.case_streams(name)
        rl = meta["workload"]["read_len"]
        s2, q2 = seq.reshape(-1, rl + 1), qual.reshape(-1, rl + 1)
        f1, f2 = wd / "a.fq", wd / "b.fq"
        O.write_fastq_fixed(f1, s2[0::2].reshape(-1), q2[0::2].reshape(-1), rl)
        O.write_fastq_fixed(f2, s2[1::2].reshape(-1), q2[1::2].reshape(-1), rl)
        return f1, f2
    return util.case_files(name)


def main():
    out = {}
    for name, thresh in CASES:
        meta = util.case_meta(name)
        with tempfile.TemporaryDirectory() as d:
            wd = Path(d)
            f1, f2 = inputs(name, wd)
            extra = ["--mode", "build"] + (["--remove_low_coverage_supernodes", str(thresh)] if thresh is not None else [])
            ref = O.run_reference(meta["k"], meta["L"], meta["W"], f1, f2, q_thresh=meta["q_thresh"], workdir=wd,
                                  extra_args=extra)
            fa = (wd / "branches.fa").read_bytes()
            recs = ref["ctx"]["records"]
            out[f"{name}:{'raw' if thresh is None else thresh}"] = {
                "branches_bytes": len(fa), "branches": fa.count(b">b_"), "branches_sha256": hashlib.sha256(fa).hexdigest(),
                "ctx_records": len(recs), "ctx_records_sha256": hashlib.sha256(recs.tobytes()).hexdigest()}
    (HERE / "branches.json").write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print(f"wrote {len(out)} cases")


if __name__ == "__main__":
    main()
