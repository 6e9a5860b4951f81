#!/bin/bash
python scripts/micro/read_bw.py
