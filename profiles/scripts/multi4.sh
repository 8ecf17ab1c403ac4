#!/bin/bash
# BASELINE configs[3]: on-ramp + roundabout maps, N = 12, 8192 envs per GPU on 4 GPUs (32768 envs in all)
TAG=onramp bash profiles/scripts/multi.sh 4 --scenario on_ramp_2_multilane --agents 12 --envs 8192 --rollout-envs 8192
TAG=roundabout bash profiles/scripts/multi.sh 4 --scenario roundabout_2 --agents 12 --envs 8192 --rollout-envs 8192
TAG=cpm bash profiles/scripts/multi.sh 4
