# builds happen in the container (nvcc cross-compiles); the GPU box only runs the binaries under build_variants/
for f in build_variants/ia_*; do [ -x "$f" ] && timeout 60 $f; done
