"""The class façade of the reference's earlier API (sydr/old/): `Acquisition` and `Tracking` objects that bind the
per-call C entry points (setSatellite, PCPS, twoCorrelationPeakComparison, generateReplica, generateCarrier,
getCorrelator, delayLockLoop, phaseLockLoop, getLoopCoefficients).  Here those nine symbols are exported by
libsydr_b200.so (csrc/legacy.cu) and run on the GPU."""
