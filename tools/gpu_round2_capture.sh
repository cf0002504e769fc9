# round-2 evidence run (one B200 through gpurun): parity suite, bench line, ncu launch list + one --set full capture, step table
python -m pytest tests -m gpu -q > gpurun_out/u24_pytest.log 2>&1; tail -4 gpurun_out/u24_pytest.log
python bench.py --steps 200 --warmup 10 > gpurun_out/u24_bench.json 2> gpurun_out/u24_bench.err; cut -c1-300 gpurun_out/u24_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager --no-parity-check --no-graph --e2e-steps 1 > gpurun_out/u24_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc|k_contract_tc|k_adam|k_ln_tanh" -s 123 -c 41 -o /tmp/r2_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cuda-eager --no-parity-check --no-graph --e2e-steps 1 > gpurun_out/u24_ncu_full.log 2>&1; tail -1 gpurun_out/u24_ncu_full.log | cut -c1-200
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null; ncu -i /tmp/r2_full.ncu-rep --page details --csv > gpurun_out/r2_full_details.csv 2>/dev/null
ncu -i /tmp/r2_full.ncu-rep --page source --csv --kernel-name regex:k_gemm_tc --launch-skip 5 --launch-count 1 > gpurun_out/r2_full_source_gemm_tc.csv 2>/dev/null
ls -la gpurun_out/ /tmp/r2_full.ncu-rep
python tools/profile_step.py > gpurun_out/r2_step_table_cuda_events.txt 2>&1; tail -3 gpurun_out/r2_step_table_cuda_events.txt
