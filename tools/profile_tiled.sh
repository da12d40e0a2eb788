# usage: tools/profile_tiled.sh <tag> [time_tiled.py args]   (run under gpurun; writes gpurun_out/prof_<tag>.ncu-rep and CSV exports)
tag=$1; shift
cmd="python tools/time_tiled.py --steps 2 $@"
$cmd > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$tag.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:assemble_tiled -c 1 -s 3 -o gpurun_out/prof_$tag -f $cmd > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv --print-source=sass > gpurun_out/sass_$tag.csv 2>/dev/null
ls -la gpurun_out/*$tag*
