"""wgpu-cpu_b200: a B200-native render-pass draw path behind wgpu-cpu's backend surface."""
