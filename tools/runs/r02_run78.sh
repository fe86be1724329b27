#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -3
