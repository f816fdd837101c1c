// TEST INFRASTRUCTURE (tests/cusim): appended to a pipeline translation unit after the kernels; the one entry point the
// stand-in cuLaunchKernel calls.  Every kernel of wgb_raster.cuh takes the draw record by value (the strip-map kernel
// one more pointer), so the parameter array of cuLaunchKernel is unpacked here by kernel name.
#pragma once
extern "C" __attribute__((visibility("default"))) void cusim_launch(void* fn, const char* name, unsigned gx, unsigned gy, unsigned gz,
                                                                    unsigned bx, unsigned by, unsigned bz, void** params) {
    const dim3 grid{gx, gy, gz}, block{bx, by, bz};
    const WgbDraw draw = *reinterpret_cast<const WgbDraw*>(params[0]);
    if (strcmp(name, "wgb_strip_map_kernel") == 0) {
        unsigned* out = *reinterpret_cast<unsigned**>(params[1]);
        auto f = reinterpret_cast<void (*)(const WgbDraw, unsigned*)>(fn);
        cusim_run_grid(grid, block, [&]() { f(draw, out); });
    } else {
        auto f = reinterpret_cast<void (*)(const WgbDraw)>(fn);
        cusim_run_grid(grid, block, [&]() { f(draw); });
    }
}
