"""videodecoder_b200 -- B200-native H.264 macroblock-reconstruction engine behind a C ABI
(include/h264r.h), with the host-side tooling around it (syntax synthesis, packing, workloads)."""
__all__ = ["abi", "build", "engine", "pack", "synth", "workload"]
