"""Re-hosted drivers of the two reference simulators that use PyLB: same command line, same
parameters, same output files -- the time loop runs on the GPU."""
