#!/bin/bash
bash profiles/scripts/variants.sh check | tail -8
bash profiles/scripts/evidence.sh ncu | tail -12
