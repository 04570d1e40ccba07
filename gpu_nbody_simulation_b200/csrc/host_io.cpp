// Host-side text formats of the reference and the canonical (reference-equivalent) node table.
#include "host_io.h"

#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>

#include "../../include/bh.h"

namespace bh {

void set_error(const char* fmt, ...);

namespace {

struct Walker {
    const HostTree& t;
    double* rows;
    int64_t cap;
    int64_t at = 0;
    FILE* f = nullptr;

    // PARTICLE_INDEX as the reference stores it (project.cu:376-378, :403, :443)
    int64_t occupant(int level, uint32_t cnt, uint32_t first) const {
        if (cnt != 1) return -1;
        uint32_t b = t.sidx[first];
        int64_t orig = t.perm.empty() ? b : t.perm[b];
        return level == t.finest ? -orig - 2 : orig;
    }

    void visit(int level, uint64_t code, double xmin, double xmax, double ymin, double ymax) {
        const uint64_t a = t.level_off[level] + code;
        const uint32_t cnt = t.count[a];
        const bool internal = cnt >= 2 && level < t.finest;
        const int64_t occ = occupant(level, cnt, t.first[a]);
        const double m = t.mass[a], cx = t.comx[a], cy = t.comy[a];
        if (rows && at < cap) {
            double* r = rows + at * 10;
            r[0] = level; r[1] = xmin; r[2] = xmax; r[3] = ymin; r[4] = ymax;
            r[5] = m; r[6] = cx; r[7] = cy; r[8] = (double)occ; r[9] = internal ? 1.0 : 0.0;
        }
        if (f) {   // project.cu:509-526, "%g" == default ostream formatting of double
            fprintf(f, "%d %g %g %g %g %g", level, xmin, xmax, ymin, ymax, m);
            if (occ != -1) {
                // Deviation from project.cu:514-518: cap-level single leaves print the
                // occupant's real position; the reference indexes positions[] with the negative code.
                uint32_t b = t.sidx[t.first[a]];
                fprintf(f, " occupantIndex=%lld occupantPos=(%g,%g)", (long long)occ, t.pos[2 * (size_t)b],
                        t.pos[2 * (size_t)b + 1]);
            } else if (m > 0) {
                fprintf(f, " occupantIndex=%lld occupantPos=(%g,%g)", (long long)occ, cx, cy);
            }
            fputc('\n', f);
        }
        ++at;
        if (!internal) return;
        const double mx = (xmin + xmax) / 2.0, my = (ymin + ymax) / 2.0;   // project.cu:417-418
        visit(level + 1, 4 * code + 0, xmin, mx, ymin, my);
        visit(level + 1, 4 * code + 1, mx, xmax, ymin, my);
        visit(level + 1, 4 * code + 2, xmin, mx, my, ymax);
        visit(level + 1, 4 * code + 3, mx, xmax, my, ymax);
    }
};

}  // namespace

int64_t canonical_rows(const HostTree& t, double* out_rows, int64_t cap_rows) {
    Walker w{t, out_rows, cap_rows};
    w.visit(0, 0, t.bounds[0], t.bounds[1], t.bounds[2], t.bounds[3]);
    return w.at;
}

int dump_quadtree_txt(const HostTree& t, const char* path) {
    FILE* f = fopen(path, "w");
    if (!f) { set_error("cannot open %s for writing", path); return BH_ERR_IO; }
    Walker w{t, nullptr, 0};
    w.f = f;
    w.visit(0, 0, t.bounds[0], t.bounds[1], t.bounds[2], t.bounds[3]);
    fclose(f);
    return BH_OK;
}

// loadSimulationDataFromText, project.cu:103-161: std::stod(line) for masses, operator>> twice for vectors.
int load_text(const char* mf, const char* pf, const char* vf, int64_t n, double* mass, double* pos, double* vel) {
    {
        std::ifstream ifs(mf);
        if (!ifs) { set_error("Failed to open file: %s", mf); return BH_ERR_IO; }
        std::string line;
        for (int64_t i = 0; i < n; ++i) {
            if (!std::getline(ifs, line)) { set_error("Not enough mass entries in file: %s", mf); return BH_ERR_IO; }
            char* end = nullptr;
            mass[i] = strtod(line.c_str(), &end);
            if (end == line.c_str()) { set_error("Failed to parse mass in file: %s", mf); return BH_ERR_IO; }
        }
    }
    auto vectors = [&](const char* name, double* out) -> int {
        std::ifstream ifs(name);
        if (!ifs) { set_error("Failed to open file: %s", name); return BH_ERR_IO; }
        std::string line;
        for (int64_t i = 0; i < n; ++i) {
            if (!std::getline(ifs, line)) { set_error("Not enough vector entries in file: %s", name); return BH_ERR_IO; }
            std::istringstream iss(line);
            for (int k = 0; k < 2; ++k)
                if (!(iss >> out[2 * i + k])) { set_error("Failed to parse vector component in file: %s", name); return BH_ERR_IO; }
        }
        return BH_OK;
    };
    int rc = vectors(pf, pos);
    if (rc != BH_OK) return rc;
    return vectors(vf, vel);
}

// savePositions, project.cu:855-863: std::to_string == "%f" for double, "%d" for int; trailing space.
int append_positions_txt(const char* path, const double* pos, int64_t n, double time, int truncate) {
    FILE* f = fopen(path, truncate ? "w" : "a");
    if (!f) { set_error("cannot open %s", path); return BH_ERR_IO; }
    for (int64_t i = 0; i < n; ++i)
        fprintf(f, "%f %lld %f %f \n", time, (long long)i, pos[2 * i], pos[2 * i + 1]);
    fclose(f);
    return BH_OK;
}

// initializeMasses / initializeVectors with saveToFile (project.cu:236-246, :268-281): `ofs << value` with the
// default ostream format, i.e. "%g" with 6 significant digits.
int write_init_files(const char* mf, const char* pf, const char* vf, int64_t n, const double* mass, const double* pos,
                     const double* vel) {
    FILE* f = fopen(mf, "w");
    if (!f) { set_error("Failed to open file for writing masses."); return BH_ERR_IO; }       // project.cu:239
    for (int64_t i = 0; i < n; ++i) fprintf(f, "%g\n", mass[i]);
    fclose(f);
    const char* names[2] = {pf, vf};
    const double* data[2] = {pos, vel};
    for (int k = 0; k < 2; ++k) {
        f = fopen(names[k], "w");
        if (!f) { set_error("Failed to open file for writing vectors."); return BH_ERR_IO; }  // project.cu:271
        for (int64_t i = 0; i < n; ++i) fprintf(f, "%g %g\n", data[k][2 * i], data[k][2 * i + 1]);
        fclose(f);
    }
    return BH_OK;
}

double round6(double v) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.6g", v);
    return strtod(buf, nullptr);
}


// ---- background trajectory writer ---------------------------------------------------------------------------------
struct FrameWriter::Impl {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    struct Job { int slot; const double* host; double time; void* ev; };
    std::deque<Job> q;
    bool busy[2] = {false, false};
    bool stop = false;
    int64_t n = 0;
    int device = 0;
};

FrameWriter::FrameWriter(const char* path, int64_t n, int device) : impl_(new Impl), f_(fopen(path, "w")) {
    impl_->n = n; impl_->device = device;
    if (!f_) { set_error("cannot open %s", path); return; }
    setvbuf((FILE*)f_, nullptr, _IOFBF, 1 << 22);
    impl_->th = std::thread([this] {
        cudaSetDevice(impl_->device);
        std::string text;
        for (;;) {
            Impl::Job job;
            {
                std::unique_lock<std::mutex> lk(impl_->mu);
                impl_->cv.wait(lk, [this] { return impl_->stop || !impl_->q.empty(); });
                if (impl_->q.empty()) return;
                job = impl_->q.front();
            }
            if (job.ev && cudaEventSynchronize((cudaEvent_t)job.ev) != cudaSuccess) failed_ = true;
            // savePositions, project.cu:855-863: std::to_string == "%f" for double, "%d" for int; trailing space
            text.clear();
            char line[128];
            for (int64_t i = 0; i < impl_->n; ++i) {
                const int len = snprintf(line, sizeof line, "%f %lld %f %f \n", job.time, (long long)i, job.host[2 * i],
                                         job.host[2 * i + 1]);
                text.append(line, (size_t)len);
                if (text.size() > (1u << 20)) { if (fwrite(text.data(), 1, text.size(), (FILE*)f_) != text.size()) failed_ = true; text.clear(); }
            }
            if (!text.empty() && fwrite(text.data(), 1, text.size(), (FILE*)f_) != text.size()) failed_ = true;
            {
                std::lock_guard<std::mutex> lk(impl_->mu);
                impl_->q.pop_front();
                impl_->busy[job.slot] = false;
            }
            impl_->cv.notify_all();
        }
    });
}

void FrameWriter::acquire(int slot) {
    std::unique_lock<std::mutex> lk(impl_->mu);
    impl_->cv.wait(lk, [&] { return !impl_->busy[slot]; });
}

void FrameWriter::submit(int slot, const double* host, double time, void* ready_event) {
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->busy[slot] = true;
        impl_->q.push_back({slot, host, time, ready_event});
    }
    impl_->cv.notify_all();
}

void FrameWriter::finish() {
    std::unique_lock<std::mutex> lk(impl_->mu);
    impl_->cv.wait(lk, [this] { return impl_->q.empty(); });
    if (f_) fflush((FILE*)f_);
}

FrameWriter::~FrameWriter() {
    if (impl_->th.joinable()) {
        finish();
        { std::lock_guard<std::mutex> lk(impl_->mu); impl_->stop = true; }
        impl_->cv.notify_all();
        impl_->th.join();
    }
    if (f_) fclose((FILE*)f_);
    delete impl_;
}

}  // namespace bh
