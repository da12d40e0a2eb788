# Round-end evidence run (1 GPU): tests, smoke, bench lines, ncu launch list, full capture of the tiled kernel, aux configs.
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tee gpurun_out/final_pytest.log | tail -1 | cut -c1-150
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null
python bench.py --path two_pass --steps 20 --no-cpu-baseline > gpurun_out/final_bench_twopass.json 2>/dev/null
python bench.py --permuted --steps 20 --no-cpu-baseline --no-e2e > gpurun_out/final_bench_permuted.json 2>/dev/null
python tools/bench_configs.py > gpurun_out/final_aux_configs.jsonl 2>/dev/null; cat gpurun_out/final_aux_configs.jsonl | cut -c1-200
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_final.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_final.log 2>&1
grep -c assemble_tiled gpurun_out/final_launches.csv
bash tools/profile_tiled.sh final --libs default --consumers 0 | tail -2
python tools/time_symbolic.py 2>/dev/null | tee gpurun_out/final_symbolic.txt
