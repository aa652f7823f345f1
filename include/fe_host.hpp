// fe_host.hpp -- ROS-free C++17 mirrors of the reference's compiled host classes, built on the C-ABI of fe_abi.h.
//
// The reference's toolchain (catkin + OpenCV C++ headers) is absent from this image, so these classes keep the
// reference's STRUCTURE -- names, threads, queues, parameters, error behaviour -- and replace every OpenCV call by
// the corresponding libfe_b200 entry point; a maintainer swaps the `#include "front_end/StereoCamera.hpp"` bodies for
// these.  Nothing here computes: all arithmetic happens in libfe_b200.so.
//
//   fe::host::StereoCamera   src/StereoCamera.cpp:5-381, include/front_end/StereoCamera.hpp:20-90
//       BufferLeft / BufferRight push image clones into queues (:43-63); processLeftImage / processRightImage threads run
//       detect + compute (:66-140), each on its own fe_ctx (the reference guards each detector with its own mutex, :76-77);
//       processStereo pops both feature queues, applies the epipolar mask |2((yL+lroi.y)-(yR+rroi.y))| <= 2, kNN-2, Lowe
//       0.8 (:143-264) and packs a StereoFrame with ROI offsets added back (:266-290).
//   fe::host::WindowMatcher  src/WindowMatcher.cpp:75-231, include/front_end/WindowMatcher.hpp
//       newStereo(): triangulate with Q (:36-51,79-85), window push / erase (:92-96), consecutive-frame search-box kNN-2 +
//       Lowe ratio (:104-231) -> currentInlierIndexes / previousInlierIndexes.  (Motion estimation :232-303 is out of scope.)
//   fe::host::LiveDetector   src/live_stereo.cpp:227-404 (stereoMatch loop): 2x3 grid FAST with the per-cell setpoint
//       controller + cornerSubPix for both eyes, descriptors, cross-check match, |dy| <= 0.7; controlDetection (:104-115).
#ifndef FE_HOST_HPP
#define FE_HOST_HPP

#include <array>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "fe_abi.h"

namespace fe {
namespace host {

struct Roi { int x = 0, y = 0, width = 0, height = 0; };          // sensor_msgs/RegionOfInterest / cv::Rect

struct Image {                                                    // cv::Mat 8UC1 clone
    int width = 0, height = 0;
    std::vector<uint8_t> data;
    Image() = default;
    Image(const uint8_t *p, int w, int h, int stride) : width(w), height(h), data((size_t)w * h) {
        for (int y = 0; y < h; ++y) std::memcpy(data.data() + (size_t)y * w, p + (size_t)y * stride, (size_t)w);
    }
};

// front_end::StereoMatch / StereoFrame of the legacy node (src/StereoCamera.cpp:266-290): one entry per accepted match
struct StereoMatch {
    float lx, ly, rx, ry;               // imageCoord with ROI offsets added back
    std::array<uint8_t, 32> ldesc, rdesc;
    float distance;
};
struct StereoFrame { std::vector<StereoMatch> matches; };

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string &m) : std::runtime_error(m), status(st) {}
};
inline void check(int st, fe_ctx *c, const char *what) {
    if (st != FE_OK) throw Error(st, std::string(what) + ": " + (fe_last_error(c) ? fe_last_error(c) : ""));
}

class Ctx {                                                       // RAII fe_ctx
  public:
    explicit Ctx(const fe_config &cfg) { check(fe_create(&cfg, &c_), nullptr, "fe_create"); }
    ~Ctx() { fe_destroy(c_); }
    Ctx(const Ctx &) = delete;
    Ctx &operator=(const Ctx &) = delete;
    fe_ctx *get() const { return c_; }

  private:
    fe_ctx *c_ = nullptr;
};

template <typename T>
class BlockingQueue {                                             // std::queue + condition_variable, as in StereoCamera.hpp:30-47
  public:
    void push(T v) {
        std::lock_guard<std::mutex> l(m_);
        q_.push(std::move(v));
        cv_.notify_one();
    }
    bool pop(T &out, const std::atomic<bool> &stop) {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return !q_.empty() || stop.load(); });
        if (q_.empty()) return false;
        out = std::move(q_.front());
        q_.pop();
        return true;
    }
    void wake() { std::lock_guard<std::mutex> l(m_); cv_.notify_all(); }

  private:
    std::queue<T> q_;
    std::mutex m_;
    std::condition_variable cv_;
};

struct Features {
    std::vector<fe_kpoint> kps;
    std::vector<uint8_t> desc;          // n x 32
};

class StereoCamera {
  public:
    using Callback = std::function<void(const StereoFrame &)>;    // stereoPub.publish(outMessage), :371

    StereoCamera(const fe_config &cfg, Roi lroi, Roi rroi, Callback publish)
        : lctx_(cfg), rctx_(cfg), mctx_(cfg), lroi_(lroi), rroi_(rroi), publish_(std::move(publish)), cap_(cfg.max_keypoints ? cfg.max_keypoints : 16384) {
        threads_.emplace_back([this] { processImage(leftImages_, leftFeatures_, lctx_); });     // :28
        threads_.emplace_back([this] { processImage(rightImages_, rightFeatures_, rctx_); });   // :29
        threads_.emplace_back([this] { processStereo(); });                                    // :30
    }
    ~StereoCamera() {
        stop_ = true;
        leftImages_.wake(); rightImages_.wake(); leftFeatures_.wake(); rightFeatures_.wake();
        for (auto &t : threads_) t.join();
    }
    void BufferLeft(const uint8_t *img, int w, int h, int stride) { leftImages_.push(Image(img, w, h, stride)); }    // :43-52
    void BufferRight(const uint8_t *img, int w, int h, int stride) { rightImages_.push(Image(img, w, h, stride)); }  // :54-63
    // front_end/setDetector (:422-521) reduced to what the hot path consumes: FAST threshold + setpoint
    int updateDetector(int threshold, int set_point) {
        int32_t sp = 0;
        check(fe_set_detection(lctx_.get(), threshold, set_point, &sp), lctx_.get(), "fe_set_detection");
        check(fe_set_detection(rctx_.get(), threshold, set_point, &sp), rctx_.get(), "fe_set_detection");
        return sp;
    }
    int framesPublished() const { return published_.load(); }

  private:
    void processImage(BlockingQueue<Image> &in, BlockingQueue<Features> &out, Ctx &ctx) {      // :66-140
        Image img;
        while (in.pop(img, stop_)) {
            Features f;
            f.kps.resize(cap_);
            int32_t n = 0;
            int st = fe_detect(ctx.get(), img.data.data(), img.width, img.height, img.width, f.kps.data(), cap_, &n);   // lDet->detect
            if (st != FE_OK && st != FE_ERR_CAPACITY) check(st, ctx.get(), "fe_detect");
            n = n < cap_ ? n : cap_;
            f.desc.resize((size_t)(n > 0 ? n : 1) * 32);
            check(fe_describe(ctx.get(), img.data.data(), img.width, img.height, img.width, f.kps.data(), &n, f.desc.data(),
                              FE_DESC_ORB256), ctx.get(), "fe_describe");                                                 // lDesc->compute
            f.kps.resize(n);
            f.desc.resize((size_t)n * 32);
            out.push(std::move(f));
        }
    }
    void processStereo() {                                                                       // :143-381
        Features l, r;
        while (leftFeatures_.pop(l, stop_) && rightFeatures_.pop(r, stop_)) {
            fe_match_cfg mc{};
            mc.ratio = 0.8; mc.mode = FE_MATCH_RATIO; mc.mask = FE_MASK_EPIPOLAR; mc.norm = FE_NORM_HAMMING;
            mc.epi_threshold = 1.0f;                      // abs(2*((yL+lroi.y)-(yR+rroi.y))) <= 2.0   (:187)
            mc.q_y_offset = (float)lroi_.y; mc.t_y_offset = (float)rroi_.y;
            std::vector<fe_match> good(l.kps.size() ? l.kps.size() : 1);
            int32_t ng = 0;
            check(fe_stereo_match(mctx_.get(), l.kps.data(), l.desc.data(), (int)l.kps.size(), r.kps.data(), r.desc.data(),
                                  (int)r.kps.size(), FE_DESC_ORB256, &mc, good.data(), (int)good.size(), &ng), mctx_.get(),
                  "fe_stereo_match");
            StereoFrame out;
            out.matches.resize(ng);
            for (int i = 0; i < ng; ++i) {                                                       // :266-290
                const fe_kpoint &lk = l.kps[good[i].queryIdx], &rk = r.kps[good[i].trainIdx];
                StereoMatch &m = out.matches[i];
                m.lx = lk.x + lroi_.x; m.ly = lk.y + lroi_.y;
                m.rx = rk.x + rroi_.x; m.ry = rk.y + rroi_.y;
                std::memcpy(m.ldesc.data(), l.desc.data() + (size_t)good[i].queryIdx * 32, 32);
                std::memcpy(m.rdesc.data(), r.desc.data() + (size_t)good[i].trainIdx * 32, 32);
                m.distance = good[i].distance;
            }
            publish_(out);
            ++published_;
        }
    }

    Ctx lctx_, rctx_, mctx_;
    Roi lroi_, rroi_;
    Callback publish_;
    int cap_;
    BlockingQueue<Image> leftImages_, rightImages_;
    BlockingQueue<Features> leftFeatures_, rightFeatures_;
    std::atomic<bool> stop_{false};
    std::atomic<int> published_{0};
    std::vector<std::thread> threads_;
};

struct InterWindowFrame { std::vector<int> currentInlierIndexes, previousInlierIndexes; };    // :227-231
struct Landmark { StereoMatch stereo; double x, y, z; };                                      // :36-51

class WindowMatcher {
  public:
    WindowMatcher(const fe_config &cfg, int nWindow, const double Q[16]) : ctx_(cfg), nWindow_(nWindow) { std::memcpy(Q_, Q, sizeof(Q_)); }
    // WindowMatcher::newStereo (:75-231); returns false when there is no previous frame yet
    bool newStereo(const StereoFrame &msg, InterWindowFrame &latestInter) {
        std::vector<Landmark> current(msg.matches.size());
        for (size_t i = 0; i < msg.matches.size(); ++i) {                                      // triangulate, :36-51
            const StereoMatch &m = msg.matches[i];
            const double in[4] = {m.lx, m.ly, (double)(m.lx - m.rx), 1.0};
            double h[4];
            for (int r = 0; r < 4; ++r) h[r] = Q_[4 * r] * in[0] + Q_[4 * r + 1] * in[1] + Q_[4 * r + 2] * in[2] + Q_[4 * r + 3] * in[3];
            current[i] = Landmark{m, h[0] / (1000 * h[3]), h[1] / (1000 * h[3]), h[2] / (1000 * h[3])};
        }
        window_.push_back(std::move(current));                                                 // :92-96
        if ((int)window_.size() >= nWindow_) window_.pop_front();
        if (window_.size() < 2) return false;
        const std::vector<Landmark> &cur = window_.back(), &prev = window_[window_.size() - 2];
        auto pack = [](const std::vector<Landmark> &v, std::vector<fe_kpoint> &k, std::vector<uint8_t> &d) {
            k.resize(v.size());
            d.resize(v.size() * 32 + 32);
            for (size_t i = 0; i < v.size(); ++i) {
                k[i] = fe_kpoint{v[i].stereo.lx, v[i].stereo.ly, 31.f, -1.f, 0.f, 0, -1};      // left feature coordinates, :112-118
                std::memcpy(d.data() + i * 32, v[i].stereo.ldesc.data(), 32);                 // left descriptors, :134-148
            }
        };
        std::vector<fe_kpoint> ck, pk;
        std::vector<uint8_t> cd, pd;
        pack(cur, ck, cd);
        pack(prev, pk, pd);
        fe_match_cfg wc{};
        wc.ratio = 0.8; wc.mode = FE_MATCH_RATIO; wc.mask = FE_MASK_WINDOW; wc.norm = FE_NORM_HAMMING; wc.win_w = 100; wc.win_h = 100;   // :32
        std::vector<fe_match> out(ck.size() ? ck.size() : 1);
        int32_t n = 0;
        check(fe_window_match(ctx_.get(), ck.data(), cd.data(), (int)ck.size(), pk.data(), pd.data(), (int)pk.size(), FE_DESC_ORB256,
                              &wc, out.data(), (int)out.size(), &n), ctx_.get(), "fe_window_match");
        latestInter.currentInlierIndexes.clear();
        latestInter.previousInlierIndexes.clear();
        for (int i = 0; i < n; ++i) {
            latestInter.currentInlierIndexes.push_back((int)out[i].queryIdx);
            latestInter.previousInlierIndexes.push_back((int)out[i].trainIdx);
        }
        return true;
    }
    const std::deque<std::vector<Landmark>> &window() const { return window_; }

  private:
    Ctx ctx_;
    int nWindow_;
    double Q_[16];
    std::deque<std::vector<Landmark>> window_;
};

class LiveDetector {
  public:
    struct Output { std::vector<fe_kpoint> left, right; std::vector<fe_match> goodMatch; };
    LiveDetector(const fe_config &cfg, Roi roi, int threshold = 15, int setPoint = 3000) : ctx_(cfg), roi_(roi), cap_(cfg.max_keypoints ? cfg.max_keypoints : 16384) {
        controlDetection(threshold, setPoint);
    }
    // fn_controlDetection (src/live_stereo.cpp:104-115); takes effect at the next frame boundary (the reference's write is
    // unsynchronised against the worker thread)
    int controlDetection(int threshold, int setPoint) {
        std::lock_guard<std::mutex> l(m_);
        setPoint_ = setPoint;
        for (int i = 0; i < 6; ++i) lThresholds_[i] = rThresholds_[i] = threshold;
        ++generation_;          // a frame in flight must not overwrite this with its stale controller state
        return setPoint_;
    }
    // cv::BriefDescriptorExtractor extractor(16) (:238): the live node describes with BRIEF-16.  Its test table is part of
    // OpenCV's sources (generated_16.i), not of this library: hand it over here (128 rows of y1, x1, y2, x2).  Without a
    // table the node falls back to rBRIEF-256 at angle -1 (same role).
    void setBriefPattern(const int8_t *tests128x4) {
        check(fe_set_brief_pattern(ctx_.get(), 16, tests128x4, 0), ctx_.get(), "fe_set_brief_pattern");
        brief_ = true;
    }
    // one iteration of stereoMatch() (:277-379) for a rectified pair
    Output process(const uint8_t *left, const uint8_t *right, int w, int h, int stride) {
        int32_t lthr[6], rthr[6];
        int sp;
        unsigned gen;
        { std::lock_guard<std::mutex> l(m_); std::memcpy(lthr, lThresholds_, sizeof(lthr)); std::memcpy(rthr, rThresholds_, sizeof(rthr)); sp = setPoint_; gen = generation_; }
        fe_grid_cfg gc{};
        gc.roi_x = roi_.x; gc.roi_y = roi_.y; gc.roi_w = roi_.width; gc.roi_h = roi_.height;    // the LEFT roi for both eyes, :272-273
        gc.rows = 2; gc.cols = 3; gc.variant = 0; gc.fast_type = FE_FAST_7_12; gc.set_point = sp; gc.subpix = 1; gc.update = 1;
        Output o;
        o.left.resize(cap_); o.right.resize(cap_);
        int32_t nl = 0, nr = 0;
        check(fe_grid_detect(ctx_.get(), left, w, h, stride, &gc, lthr, o.left.data(), cap_, &nl, nullptr), ctx_.get(), "fe_grid_detect");
        check(fe_grid_detect(ctx_.get(), right, w, h, stride, &gc, rthr, o.right.data(), cap_, &nr, nullptr), ctx_.get(), "fe_grid_detect");
        {   // write the controller's step back unless controlDetection() landed meanwhile: its values win and take effect
            // at the next frame boundary (in the reference the service write persists because the worker updates in place)
            std::lock_guard<std::mutex> l(m_);
            if (gen == generation_) { std::memcpy(lThresholds_, lthr, sizeof(lthr)); std::memcpy(rThresholds_, rthr, sizeof(rthr)); }
        }
        std::vector<uint8_t> ld((size_t)(nl > 0 ? nl : 1) * 32), rd((size_t)(nr > 0 ? nr : 1) * 32);
        // extractor.compute(currentLeft, leftKP, lDescriptor) (:359-360): BRIEF-16 when its table was supplied
        const int32_t kind = brief_ ? FE_DESC_BRIEF16 : FE_DESC_ORB256;
        check(fe_describe(ctx_.get(), left, w, h, stride, o.left.data(), &nl, ld.data(), kind), ctx_.get(), "fe_describe");
        check(fe_describe(ctx_.get(), right, w, h, stride, o.right.data(), &nr, rd.data(), kind), ctx_.get(), "fe_describe");
        o.left.resize(nl); o.right.resize(nr);
        fe_match_cfg mc{};
        mc.ratio = 0.8; mc.mode = FE_MATCH_CROSSCHECK; mc.mask = FE_MASK_NONE; mc.norm = FE_NORM_HAMMING; mc.max_dy = 0.7f;       // :364-377
        o.goodMatch.resize(nl > 0 ? nl : 1);
        int32_t ng = 0;
        check(fe_stereo_match(ctx_.get(), o.left.data(), ld.data(), nl, o.right.data(), rd.data(), nr, kind, &mc,
                              o.goodMatch.data(), (int)o.goodMatch.size(), &ng), ctx_.get(), "fe_stereo_match");
        o.goodMatch.resize(ng);
        return o;
    }
    const int32_t *leftThresholds() const { return lThresholds_; }

  private:
    Ctx ctx_;
    Roi roi_;
    int cap_;
    std::mutex m_;
    int setPoint_ = 3000;
    bool brief_ = false;
    unsigned generation_ = 0;
    int32_t lThresholds_[6], rThresholds_[6];
};

}  // namespace host
}  // namespace fe

#endif  // FE_HOST_HPP
