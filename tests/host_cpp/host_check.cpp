// host_check -- drives the C++ host mirrors of include/fe_host.hpp (StereoCamera, WindowMatcher, LiveDetector) over a
// sequence of stereo frames read from a file and writes what they publish to another file; tests/test_host_cpp.py
// compares that with the oracle.  Links ONLY against libfe_b200.so (the C-ABI); exits 3 with the library's message when
// there is no CUDA device (there is no CPU fallback).
//
//   in : int32 {W, H, F, n_features, fast_threshold, roi_x, roi_y, roi_w, roi_h, set_point, live_threshold}
//        F x (left W*H u8, right W*H u8), Q 16 x f64
//   out: per frame  int32 n, n x {lx, ly, rx, ry, distance} f32, n x 32 u8 (left descr), n x 32 u8 (right descr)
//        per frame>0 int32 n, n x int32 current index, n x int32 previous index
//        per frame  (LiveDetector) 6 x int32 left thresholds after the frame, int32 nl, nr, ng,
//                   nl x fe_kpoint, nr x fe_kpoint, ng x fe_match
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "fe_host.hpp"

template <typename T>
static void put(FILE *f, const T *p, size_t n) {
    if (n && std::fwrite(p, sizeof(T), n, f) != n) { std::perror("write"); std::exit(2); }
}

int main(int argc, char **argv) {
    if (argc != 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE *in = std::fopen(argv[1], "rb");
    if (!in) { std::perror(argv[1]); return 2; }
    int32_t hd[11];
    if (std::fread(hd, sizeof(int32_t), 11, in) != 11) return 2;
    const int W = hd[0], H = hd[1], F = hd[2];
    std::vector<std::vector<uint8_t>> L(F), R(F);
    for (int f = 0; f < F; ++f) {
        L[f].resize((size_t)W * H); R[f].resize((size_t)W * H);
        if (std::fread(L[f].data(), 1, L[f].size(), in) != L[f].size()) return 2;
        if (std::fread(R[f].data(), 1, R[f].size(), in) != R[f].size()) return 2;
    }
    double Q[16];
    if (std::fread(Q, sizeof(double), 16, in) != 16) return 2;
    std::fclose(in);

    fe_config cfg;
    fe_default_config(&cfg);             // cv::ORB defaults: FAST-9_16 + NMS, edge 31, intensity-centroid orientation
    cfg.max_width = W; cfg.max_height = H; cfg.max_images = 6; cfg.max_keypoints = 8192;
    cfg.n_features = hd[3]; cfg.fast_threshold = hd[4]; cfg.orientation = 1;
    try {
        FILE *out = std::fopen(argv[2], "wb");
        if (!out) { std::perror(argv[2]); return 2; }
        std::vector<fe::host::StereoFrame> frames;
        std::mutex fm;
        {
            // legacy threaded node: three workers, images buffered from the "subscriber" thread
            fe::host::StereoCamera cam(cfg, fe::host::Roi{}, fe::host::Roi{}, [&](const fe::host::StereoFrame &s) {
                std::lock_guard<std::mutex> l(fm);
                frames.push_back(s);
            });
            for (int f = 0; f < F; ++f) {
                cam.BufferLeft(L[f].data(), W, H, W);
                cam.BufferRight(R[f].data(), W, H, W);
            }
            for (int spin = 0; cam.framesPublished() < F; ++spin) {
                if (spin > 60000) { std::fprintf(stderr, "StereoCamera did not publish %d frames\n", F); return 4; }
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
        }
        for (const auto &s : frames) {
            int32_t n = (int32_t)s.matches.size();
            put(out, &n, 1);
            for (const auto &m : s.matches) { const float v[5] = {m.lx, m.ly, m.rx, m.ry, m.distance}; put(out, v, 5); }
            for (const auto &m : s.matches) put(out, m.ldesc.data(), 32);
            for (const auto &m : s.matches) put(out, m.rdesc.data(), 32);
        }
        fe::host::WindowMatcher win(cfg, 10, Q);
        for (int f = 0; f < F; ++f) {
            fe::host::InterWindowFrame iw;
            const bool have = win.newStereo(frames[f], iw);
            if (have != (f > 0)) { std::fprintf(stderr, "WindowMatcher: unexpected state at frame %d\n", f); return 4; }
            if (!have) continue;
            int32_t n = (int32_t)iw.currentInlierIndexes.size();
            put(out, &n, 1);
            put(out, iw.currentInlierIndexes.data(), n);
            put(out, iw.previousInlierIndexes.data(), n);
        }
        fe::host::LiveDetector live(cfg, fe::host::Roi{hd[5], hd[6], hd[7], hd[8]}, hd[10], hd[9]);
        for (int f = 0; f < F; ++f) {
            auto o = live.process(L[f].data(), R[f].data(), W, H, W);
            put(out, live.leftThresholds(), 6);
            const int32_t c[3] = {(int32_t)o.left.size(), (int32_t)o.right.size(), (int32_t)o.goodMatch.size()};
            put(out, c, 3);
            put(out, o.left.data(), o.left.size());
            put(out, o.right.data(), o.right.size());
            put(out, o.goodMatch.data(), o.goodMatch.size());
        }
        std::fclose(out);
    } catch (const fe::host::Error &e) {
        std::fprintf(stderr, "host_check: %s (status %d)\n", e.what(), e.status);
        return e.status == FE_ERR_NO_DEVICE ? 3 : 4;
    }
    return 0;
}
