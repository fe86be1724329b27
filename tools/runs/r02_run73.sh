#!/bin/bash
python tools/cold_call_phases.py 4096 12 alternate 2>&1 | grep "cold call"
