#!/bin/bash
python tools/dropin_slots.py 4096 2>&1 | grep round
