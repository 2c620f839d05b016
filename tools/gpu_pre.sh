timeout 600 python -m pytest tests -m gpu -q -k "face_arrays or restart_from_reference" 2>&1 | tail -4
for mode in host graph; do timeout 300 python tools/ab_bench.py --degree 7 --elems 32 --mode $mode 2>&1 | cut -c1-300; done
timeout 300 python tools/ab_bench.py --degree 5 --elems 32 --mode graph 2>&1 | cut -c1-300
