#!/bin/bash
./openmm_rigidbody_plugin_b200/lib_exp/tile_stream2 > gpurun_out/r02_tile_stream2.txt 2>&1; cat gpurun_out/r02_tile_stream2.txt
./openmm_rigidbody_plugin_b200/lib_exp/tile_stream2 | tail -4
