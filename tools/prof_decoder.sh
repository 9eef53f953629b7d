# one ncu --set full capture of the decoder kernel (decode of 256 Ki rows: same per-tile behaviour, shorter replays)
T=${1:-r2}
export TD_ROWS=${TD_ROWS:-262144}
timeout 200 python tools/time_decoder.py child > gpurun_out/${T}_dec_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decoder_tc -s 1 -c 1 -o gpurun_out/prof_${T}_decoder -f python tools/time_decoder.py child > gpurun_out/${T}_ncu_dec.log 2>&1
tail -3 gpurun_out/${T}_dec_plain.log; tail -3 gpurun_out/${T}_ncu_dec.log; ls -la gpurun_out/prof_${T}_decoder.ncu-rep
