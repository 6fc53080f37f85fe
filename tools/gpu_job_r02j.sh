#!/bin/bash
# one GPU call: tests, the bench lines, the ncu launch list and the two full captures kept under profiles/
O=gpurun_out
timeout 800 python -m pytest tests -m gpu -q > $O/r02j_pytest.log 2>&1; tail -4 $O/r02j_pytest.log
timeout 280 python bench.py > $O/r02j_bench.json 2> $O/r02j_bench.err; tail -c 300 $O/r02j_bench.err; cut -c1-300 $O/r02j_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02j_bench_reference.json 2> $O/r02j_bench_reference.err; cut -c1-400 $O/r02j_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02j_launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu > $O/r02j_ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cvt_tree_kernel -c 2 -o $O/r02j_tree -f python tools/bench_extras.py --what tree5 > $O/r02j_ncu_tree.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:insert_kernel -s 4 -c 1 -o $O/r02j_grid_insert -f python tools/prof_grid.py 6 > $O/r02j_ncu_grid.log 2>&1
timeout 120 python tools/grid_insert_probe.py > $O/r02j_grid_probe.txt 2>&1; cat $O/r02j_grid_probe.txt
timeout 200 python tools/small_add_probe.py > $O/r02j_small_add.txt 2>&1; head -c 1500 $O/r02j_small_add.txt
ls -la $O/r02j_*
