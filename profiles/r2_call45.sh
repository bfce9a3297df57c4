#!/bin/bash
cd "$(dirname "$0")/.."
bash profiles/r2_call5_light_sample.sh
bash profiles/r2_call4_full_bench.sh
