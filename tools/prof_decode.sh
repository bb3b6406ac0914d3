# one ncu --set full capture of each batched-decode kernel (config 2), all lists of the synthetic block_optpfor index
mkdir -p gpurun_out/final
ncu --set full --clock-control none --import-source on -k regex:decode_full_blocks_kernel -s 1 -c 1 -o gpurun_out/final/decode_full_prof -f python tools/prof_decode.py > gpurun_out/final/ncu_decode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_serial_blocks_kernel -s 1 -c 1 -o gpurun_out/final/decode_serial_prof -f python tools/prof_decode.py >> gpurun_out/final/ncu_decode.log 2>&1
tail -2 gpurun_out/final/ncu_decode.log
