python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/microbench.py decode 2>/dev/null | tee gpurun_out/micro_decode.jsonl | cut -c1-330
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/micro_decode_launches.csv python tools/microbench.py decode --steps 1 --warmup 1 > /dev/null 2>&1
awk -F'","' 'NR>4{print $5, $NF}' gpurun_out/micro_decode_launches.csv
