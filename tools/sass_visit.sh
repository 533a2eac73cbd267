#!/bin/bash
# CPU-side check used before GPU time is spent (profiles/README.md "Alu-pipe relief"): compile rrtb_render.cu for sm_100a with
# the given -D switches, dump the headline render kernel's SASS and count the instructions of the STEP loop (two node visits +
# the continue-vote), by opcode and for the opcodes that run on the half-rate alu pipe.
# usage: tools/sass_visit.sh <name> [-D...]     -> /tmp/rrtb_sass/<name>.sass
name=${1:-head}; shift
out=/tmp/rrtb_sass; mkdir -p $out
cd "$(dirname "$0")/../rrt_b200/csrc"
K=_ZN4rrtb13k_render_poolILb0ELi2ELb0ENS_7PathF32ELb0EEEvNS_10RenderArgsE
nvcc -ccbin /usr/bin/g++ -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xptxas -v "$@" \
     -cubin rrtb_render.cu -o $out/$name.cubin 2> $out/$name.log || { tail -20 $out/$name.log; exit 1; }
grep -A2 "Compiling entry function '$K" $out/$name.log | tail -2
cuobjdump -sass -fun $K $out/$name.cubin | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+//; s/\s+\/\* 0x[0-9a-f]+ \*\/$//' > $out/$name.sass
python3 - $out/$name.sass <<'P'
import collections, re, sys
L = [l.rstrip() for l in open(sys.argv[1])]
print("kernel:", len(L), "instructions")
idx = [i for i, l in enumerate(L) if "LDG.E.ENL2.256" in l]  # three 256-bit loads per node visit
h = idx[0]
while "BSSY" not in L[h]:
    h -= 1
h -= 1  # the loop head: the `cur >= 0` test in front of the first visit
e = idx[5]
while "VOTE.ANY" not in L[e]:
    e += 1
while "BRA" not in L[e]:
    e += 1  # the back edge after the continue-vote
seg = L[h:e + 1]
c = collections.Counter()
for l in seg:
    m = re.match(r"/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    c[m.group(2).split(".")[0]] += 1
alu = {"FMNMX", "FMNMX3", "FSETP", "LOP3", "SEL", "FSEL", "VIMNMX", "VIMNMX3", "ISETP", "SHF", "VIADD", "IADD3", "MOV", "PRMT", "LEA",
       "VOTE", "POPC", "FSET"}
print("STEP loop (2 visits + vote):", len(seg), "instructions,", sum(v for k, v in c.items() if k in alu), "on the alu pipe")
print(sorted(c.items(), key=lambda x: -x[1]))
P
