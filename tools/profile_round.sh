# The round's profile captures (one GPU): launch list of the bench, ncu --set full of one substep (cube, dam slab),
# in-graph timelines, and the N=1 bench line. Outputs under gpurun_out/.
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_1M_cube.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_b_under_ncu.log 2>&1
B200MPM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"k_touch|k_block_prepare|k_scatter|k_p2g|k_g2p" -s 50 -c 5 -o gpurun_out/r02_full_1M_cube -f python tools/run_config.py cube1m 1 > /dev/null 2>&1
B200MPM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"k_touch|k_block_prepare|k_scatter|k_p2g|k_g2p" -s 50 -c 5 -o gpurun_out/r02_full_dam2m -f python tools/run_config.py dam2m 1 > /dev/null 2>&1
python tools/timeline.py cube1m > gpurun_out/r02_timeline_cube.txt 2>&1
python tools/timeline.py dam2m > gpurun_out/r02_timeline_dam.txt 2>&1
timeout 900 python bench.py 2>&1 | grep "^{" | tail -1 > gpurun_out/r02_bench_n1_final.json
ls -la gpurun_out | tail -8
