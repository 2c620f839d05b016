#!/bin/bash
# registers / spills / shared memory per kernel instance of one degree: tools/ptxas_res.sh <N> [extra nvcc flags]
N=$1; shift
cd "$(dirname "$0")/../galaexi_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xptxas -v "$@" -DDGX_N=$N -c dgx_inst.cu -o /tmp/ptxas_N$N.o 2>&1 \
 | grep -E "Compiling entry|Used|spill" | paste - - - \
 | sed -E 's/ptxas info\s*: //g; s/Compiling entry function .(_Z[A-Za-z0-9_]*). for .sm_100a./\1/' \
 | while read -r line; do
     nm=$(echo "$line" | awk '{print $1}' | c++filt | sed -E 's/\(dgx::KParams.*//; s/void dgx:://')
     echo "$nm :: $(echo "$line" | cut -d' ' -f2-)"
   done | sort
