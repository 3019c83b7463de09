#!/bin/bash
for p in 4096 8192 16384 32768 65536; do python tools/train_chunk_probe.py $p 40960 2>&1 | tail -2; done
echo "---- 16384 points per crop, workspace 12 GB / 80 GB"
python tools/train_chunk_probe.py 16384 12288 2>&1 | tail -2
python tools/train_chunk_probe.py 16384 81920 2>&1 | tail -2
