set -x
python tools/trace_late_sweep.py 30 256 12 notrace > gpurun_out/p_plain.out 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:jacobi_gram_kernel --launch-skip 640 -c 1 -o gpurun_out/r02_jacobi_gram -f python tools/trace_late_sweep.py 30 256 12 notrace > gpurun_out/p_ncu1.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:gs_panel_kernel --launch-skip 3000 -c 1 -o gpurun_out/r02_gs_panel -f python tools/trace_late_sweep.py 30 256 12 notrace > gpurun_out/p_ncu2.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:gs_update_kernel --launch-skip 3000 -c 1 -o gpurun_out/r02_gs_update -f python tools/trace_late_sweep.py 30 256 12 notrace > gpurun_out/p_ncu3.out 2>&1
python bench.py --steps 3 --warmup 14 > gpurun_out/p_bench_plain.json 2> gpurun_out/p_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 112000 -c 9000 --csv --log-file gpurun_out/r02_launches_late.csv python bench.py --steps 3 --warmup 14 > gpurun_out/p_bench_ncu.out 2>&1
tail -2 gpurun_out/p_ncu1.out gpurun_out/p_ncu2.out gpurun_out/p_ncu3.out; wc -l gpurun_out/r02_launches_late.csv; ls -la gpurun_out/*.ncu-rep
