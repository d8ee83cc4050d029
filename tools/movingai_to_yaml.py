#!/usr/bin/env python3
"""movingai.com MAPF benchmark (.scen + .map) -> the YAML files the cbs / ecbs CLIs read.
Same arguments and output names as the reference's example/standard_benchmark_converter.py
(<output_prefix>_<k>_agents.yaml for k = 10, 20, ...); unlike that script, map obstacles are kept
(see libmultirobotplanning_b200/instances.py)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmultirobotplanning_b200 import instances  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("scenario", help=".scen Scenario file")
ap.add_argument("map", help=".map Map file")
ap.add_argument("output_prefix", help=".yaml Output file prefix")
ap.add_argument("--min-agents", type=int, default=10)
ap.add_argument("--agent-step", type=int, default=10)
args = ap.parse_args()
for path in (args.scenario, args.map):
    if not os.path.isfile(path):
        print("%s not found!" % path)
        sys.exit(-1)
k = args.min_agents
for inst in instances.movingai_instances(args.scenario, args.map, args.min_agents, args.agent_step):
    name = "%s_%d_agents.yaml" % (args.output_prefix, inst.n_agents)
    print("Generating", name)
    instances.save_yaml(inst, name)
