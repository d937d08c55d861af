#!/bin/bash
for v in 1 2 3 4 5; do timeout 60 ./scripts/umma_probe $v 2>&1 | tail -6; done
