#!/bin/bash
# Diagnostic: the default bench with progress markers on stderr and a faulthandler dump if it is
# still running after 115 s (one earlier run on a slow box ended without a JSON line).
mkdir -p gpurun_out
timeout 135 python -X faulthandler -c "import faulthandler,sys,runpy; faulthandler.dump_traceback_later(115, exit=False); sys.argv=['bench.py']; runpy.run_path('bench.py', run_name='__main__')" > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
echo "rc=$?"; tail -c 300 gpurun_out/r02j_bench.json; echo; grep "bench " gpurun_out/r02j_bench.err | tail -14; grep -v "bench " gpurun_out/r02j_bench.err | tail -25
