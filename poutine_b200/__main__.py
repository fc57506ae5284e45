"""`python -m poutine_b200 -u -f anc.fasta -t anc.newick -p pheno.tsv -m sites.map -d out_dir` — the reference's own
flags (picocli options of Homoplasy_Counter, HC.java:59-154) for the --use-precomputed-anc-recon mode, JVM-less:
loaders + packer + writer from hostio.py around the GPU hot path.  The non -u mode (running treetime) is the Java
host's business and is not offered here."""
import argparse
import os
import sys
import time

from . import hostio
from .counter import NoInformativeSites, PoutineError


def main(argv=None):
    ap = argparse.ArgumentParser(prog="poutine_b200")
    ap.add_argument("-f", "--fasta", required=True, help="ancestral multi-fasta (variable sites only, every tree node)")
    ap.add_argument("-p", "--phenos", required=True)
    ap.add_argument("-t", "--tree", required=True, help="ancestral Newick with internal labels, or treetime's annotated_tree.nexus")
    ap.add_argument("-m", "--map", required=True)
    ap.add_argument("-r", "--replicates", type=int, default=100000)          # HC.java:89
    ap.add_argument("-c", "--min_hcount", type=int, default=1)               # HC.java:98
    ap.add_argument("-u", "--use-precomputed-anc-recon", action="store_true")
    ap.add_argument("-T", "--threads", type=int, default=0, help="host packer threads (the hot path itself runs on the GPU)")
    ap.add_argument("-d", "--out-dir", default=None)
    ap.add_argument("-o", "--out", default=None)
    ap.add_argument("-X", "--force-overwrite", action="store_true")
    ap.add_argument("--seed", type=int, default=None, help="Philox seed (default: from the clock, like the reference's unseeded RNG)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--extras", action="store_true", help="also write <out>.extras: exact point-wise null + two-sided Fisher "
                                                          "(not reference outputs; the three reference files are unchanged)")
    a = ap.parse_args(argv)
    if not a.use_precomputed_anc_recon:
        ap.error("only --use-precomputed-anc-recon (-u) mode is supported: ancestral reconstruction stays with the Java host / treetime")
    stamp = time.strftime("%Y_%m_%d_%H_%M_%S")
    out_dir = a.out_dir or ("poutine_session_" + stamp)
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, a.out or ("poutine_session_" + stamp + ".out"))
    if os.path.exists(out) and not a.force_overwrite:
        print("output file %s exists (use -X to overwrite)" % out, file=sys.stderr)
        return 1
    seed = a.seed if a.seed is not None else time.time_ns() & 0xFFFFFFFFFFFF
    t0 = time.time()
    try:
        res = hostio.run_homoplasy_counter(a.fasta, a.tree, a.phenos, a.map, out, num_replicates=a.replicates, min_hcount=a.min_hcount,
                                           seed=seed, device=a.device, pack_threads=a.threads, extras=a.extras)
    except NoInformativeSites:
        print("min_hcount = %d is set too high and has filtered out all segregating sites" % a.min_hcount)     # HC.java:3866-3869
        return 255
    except (PoutineError, hostio.DataFormatError, OSError) as e:
        print("error: %s" % e, file=sys.stderr)
        return 255
    c = res["counts"]
    print("# bi-allelic sites = %d, monomorphic = %d, tri-allelic = %d, quad-allelic = %d" % (c["bi"], c["mono"], c["tri"], c["quad"]))
    print("# homoplasically informative sites = %d" % res["n_informative"])
    print("results: %s (+ .sorted_by_a1_maxT, .sorted_by_a2_maxT); wall clock %.2f s" % (out, time.time() - t0))
    return 0


if __name__ == "__main__":
    sys.exit(main())
