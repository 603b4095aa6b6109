for mb in 4 5; do cp devlibs/lib_mb$mb.so glmmadaptive_b200/libglmma.so; echo "== minblocks $mb"; bash devrun.sh; done
