#!/bin/bash
# Round 2c GPU session, ONE gpurun call (most important first; every phase under its own timeout, logs under gpurun_out/):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/gpu_r02c.sh'
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r02c_clocks.csv &
SMI=$!
t0=$(date +%s)
echo "== parity + sweep"; timeout 560 python scripts/gpu_r02c.py parity sweep --patch > gpurun_out/r02c_session.log 2>&1; echo "rc=$? t=$(( $(date +%s) - t0 ))"
tail -3 gpurun_out/r02c_lean_parity.txt; grep -E "best:|patched" gpurun_out/r02c_lean_sweep.txt
echo "== tests on the patched table"
timeout 260 python -m pytest -x -q -m gpu "tests/test_gpu_parity.py::test_full_size_vs_reference_cpu_backend[1-3]" tests/test_gpu_parity.py::test_streamed_host_buffer_apply_is_bitwise_equal \
  tests/test_gpu_parity.py::test_lean_kernel_vector_qdata_loads_with_any_alignment tests/test_gpu_parity.py::test_deterministic_scatter_is_bitwise_reproducible_and_equals_serial_order \
  tests/test_gpu_libceed_backend.py::test_host_resident_apply_streams_and_matches tests/test_gpu_libceed_backend.py::test_operator_through_libceed_matches_cpu_reference \
  > gpurun_out/r02c_pytest_subset.txt 2>&1; echo "rc=$? t=$(( $(date +%s) - t0 ))"; tail -3 gpurun_out/r02c_pytest_subset.txt
echo "== ncu --set full of the tuned BP1 p=3 kernel + finalize"
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:b200_operator|k_halo_finalize' -s 6 -c 2 -f -o gpurun_out/prof_bp1p3_r02c python scripts/gpu_one.py bp1p3 > gpurun_out/r02c_ncu_full.log 2>&1
echo "rc=$? t=$(( $(date +%s) - t0 ))"
python scripts/ncu_summary.py gpurun_out/prof_bp1p3_r02c.ncu-rep > gpurun_out/r02c_ncu_lean_summary.txt 2>&1
python scripts/ncu_bounds.py "BP1 p=3 tuned lean kernel=gpurun_out/prof_bp1p3_r02c.ncu-rep" > gpurun_out/r02c_ncu_bounds.txt 2>&1
echo "== launch list of the bench command"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches_bp1p3.csv python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu-baseline --no-scaling-case --no-e2e-pipeline > gpurun_out/r02c_launch_bench.log 2>&1
echo "rc=$? t=$(( $(date +%s) - t0 ))"
echo "== bench line"
timeout 600 python bench.py > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err; echo "rc=$? t=$(( $(date +%s) - t0 ))"
kill $SMI
head -c 600 gpurun_out/r02c_bench_1gpu.json
