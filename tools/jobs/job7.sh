#!/bin/bash
timeout 300 python tools/bench_callers.py ppo --no-validate > gpurun_out/ppo_noval.log 2>&1; tail -1 gpurun_out/ppo_noval.log | cut -c1-700
timeout 300 python tools/bench_callers.py ppo --graph > gpurun_out/ppo_graph.log 2>&1; tail -1 gpurun_out/ppo_graph.log | cut -c1-700
