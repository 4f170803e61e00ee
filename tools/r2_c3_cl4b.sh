#!/bin/bash
mkdir -p gpurun_out
( GB200_LSTM_REC_CL=4 timeout 300 python bench.py --config c3 --steps 1 --warmup 0 --cpu-seconds 0 --no-extras --streams 8 --machines 16 ) > gpurun_out/r2i_c3_cl4_small.json 2> gpurun_out/r2i_c3_cl4_small.err; echo "small rc=$?"; tail -c 300 gpurun_out/r2i_c3_cl4_small.json; tail -2 gpurun_out/r2i_c3_cl4_small.err
( GB200_LSTM_REC_CL=4 timeout 600 python bench.py --config c3 --steps 1 --warmup 0 --cpu-seconds 0 --no-extras --streams 32 ) > gpurun_out/r2i_c3_cl4b.json 2> gpurun_out/r2i_c3_cl4b.err; echo "full rc=$?"; tail -c 400 gpurun_out/r2i_c3_cl4b.json; tail -3 gpurun_out/r2i_c3_cl4b.err
