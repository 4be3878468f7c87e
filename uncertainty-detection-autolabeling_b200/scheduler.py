"""Image-parallel batch scheduler - placeholder until the multi-GPU path lands."""
