#!/bin/bash
tools/ab_run.sh 3 pe_now pe_ahead -- python tools/pe_bench.py 256 30
