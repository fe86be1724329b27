#!/bin/bash
VITB200_DEBUG_TEARDOWN=1 python tools/cold_call_phases.py 4096 8 2>&1 | grep -E "cold call|tear-down:"
