"""Reference-shaped GPU baseline (test / measurement infrastructure; see ref_rasterizer.cu)."""
