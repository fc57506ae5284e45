#!/usr/bin/env python
"""AddressSanitizer + UBSan pass over the host-side C++ that parses untrusted text: the native host's loaders
(poutine_b200/host/hostio.cpp through the binary's --x-* probes) and the library's FASTA indexer / tile packer / plane
compactor (csrc/fasta.cpp, csrc/packer.cpp through tools/asan_fasta_main.cpp), on every input of
tests/golden/{loaders,hostside}_refbytecode.json (hand-written corner cases + seeded fuzz).  No GPU involved.
usage: python tools/sanitize_host.py   -> prints the number of runs and of sanitizer reports (expected: 0)"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = tempfile.mkdtemp(prefix="pb_asan_")
flags = ["-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I" + os.path.join(ROOT, "include")]
host = os.path.join(d, "host_asan")
libdir = os.path.join(ROOT, "poutine_b200")
subprocess.run(["g++", *flags, os.path.join(libdir, "host", "hostio.cpp"), os.path.join(libdir, "host", "poutine_main.cpp"), "-o", host,
                "-L" + libdir, "-lpoutine_b200", "-Wl,-rpath," + libdir, "-lpthread"], check=True)
fasta = os.path.join(d, "fasta_asan")
subprocess.run(["g++", *flags, "-mavx2", "-I" + os.path.join(libdir, "csrc"), os.path.join(ROOT, "tools", "asan_fasta_main.cpp"),
                os.path.join(libdir, "csrc", "fasta.cpp"), os.path.join(libdir, "csrc", "packer.cpp"), "-o", fasta, "-lpthread"], check=True)
env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
runs = problems = 0


def run(cmd, raw, name, ok=(0, 255)):
    global runs, problems
    p = os.path.join(d, name)
    with open(p, "wb") as f:
        f.write(raw)
    r = subprocess.run(cmd + [p], capture_output=True, env=env, timeout=120)
    runs += 1
    if r.returncode not in ok or b"runtime error" in r.stderr or b"Sanitizer" in r.stderr:
        problems += 1
        print("PROBLEM", cmd[1:], raw[:60], r.returncode, r.stderr[-500:].decode(errors="replace"))


g = json.load(open(os.path.join(ROOT, "tests", "golden", "loaders_refbytecode.json")))
h = json.load(open(os.path.join(ROOT, "tests", "golden", "hostside_refbytecode.json")))
for c in g["newick"]:
    run([host, "--x-parse-tree"], c["text"].encode("utf-8"), "t.nwk")
for c in h["nexus"]:
    run([host, "--x-nexus-to-newick"], c["raw_latin1"].encode("latin-1"), "annotated_tree.nexus")
for c in h["phenos"]:
    run([host, "--x-read-phenos"], c["raw_latin1"].encode("latin-1"), "p.tsv")
for c in h["map"]:
    run([host, "--x-read-map"], c["raw_latin1"].encode("latin-1"), "m.map")
env["ASAN_OPTIONS"] = "detect_leaks=1"
for c in g["fasta"]:
    for threads in ("1", "3"):
        p = os.path.join(d, "a.fasta")
        with open(p, "wb") as f:
            f.write(c["raw_latin1"].encode("latin-1"))
        r = subprocess.run([fasta, p, threads], capture_output=True, env=env, timeout=120)
        runs += 1
        if r.returncode != 0 or b"runtime error" in r.stderr or b"Sanitizer" in r.stderr:
            problems += 1
            print("PROBLEM fasta", repr(c["raw_latin1"][:50]), r.returncode, r.stderr[-500:].decode(errors="replace"))
print("sanitizer runs %d, reports %d" % (runs, problems))
sys.exit(1 if problems else 0)
