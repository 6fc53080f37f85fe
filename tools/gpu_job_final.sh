#!/bin/bash
# final evidence of the round: tests, smoke, both bench arms, launch list
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02z_pytest.log 2>&1; tail -3 $O/r02z_pytest.log
timeout 300 python __graft_entry__.py > $O/r02z_smoke.log 2>&1; tail -1 $O/r02z_smoke.log
timeout 300 python bench.py > $O/r02z_bench.json 2> $O/r02z_bench.err; cut -c1-250 $O/r02z_bench.json
timeout 300 python bench.py --impl reference > $O/r02z_bench_reference.json 2> $O/r02z_bench_reference.err; cut -c1-250 $O/r02z_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02z_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu > $O/r02z_ncu_launches.log 2>&1
python tools/summarize_ncu.py launches $O/r02z_launches.csv $O/r02z_launches.txt; head -8 $O/r02z_launches.txt
