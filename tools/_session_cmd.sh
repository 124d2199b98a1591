RPOOL_VARIANT_CFGS="1 13 0 3" bash tools/gpu_variants.sh r02h "" "forward_order=2" "forward_order=2,cta_threads=256" "cta_threads=256" "cta_threads=192"
