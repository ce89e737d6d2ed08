// See euroc_ingest.hpp.  File formats: ASL "data.csv" tables (first line is a header, '#' starts a
// comment, timestamps in integer nanoseconds), PNG (ISO/IEC 15948) restricted to colour type 0, bit depth 8,
// no interlace -- what the EuRoC cameras record.
#include "euroc_ingest.hpp"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

namespace aria::adapters::b200 {

namespace {

bool set_err(std::string* err, const std::string& what) { if (err) *err = what; return false; }

uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

bool read_file(const std::string& path, std::vector<uint8_t>& out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = n >= 0 && fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}

bool file_exists(const std::string& p) { std::ifstream f(p); return f.good(); }

// One table: numeric columns of every data row; `text_col` (if >= 0) is returned as a trimmed string instead.
struct CsvRow { std::vector<double> num; std::string text; };

bool read_csv(const std::string& path, size_t min_cols, int text_col, std::vector<CsvRow>& rows)
{
    std::ifstream in(path);
    if (!in.is_open()) return false;
    std::string line;
    std::getline(in, line);                                   // header
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#') continue;
        CsvRow r;
        size_t pos = 0, col = 0;
        bool bad = false;
        while (pos <= line.size()) {
            size_t comma = line.find(',', pos);
            if (comma == std::string::npos) comma = line.size();
            std::string cell = line.substr(pos, comma - pos);
            const size_t a = cell.find_first_not_of(" \t\r\n"), b = cell.find_last_not_of(" \t\r\n");
            cell = a == std::string::npos ? std::string() : cell.substr(a, b - a + 1);
            if ((int)col == text_col) r.text = cell;
            else {
                char* end = nullptr;
                const double v = strtod(cell.c_str(), &end);
                if (end == cell.c_str()) { bad = true; break; }
                r.num.push_back(v);
            }
            ++col;
            pos = comma + 1;
        }
        if (!bad && col >= min_cols) rows.push_back(std::move(r));
    }
    return true;
}

bool yaml_list(const std::string& line, const char* key, double* out, int n)
{
    if (line.find(key) == std::string::npos) return false;
    const size_t a = line.find('['), b = line.find(']');
    if (a == std::string::npos || b == std::string::npos || b < a) return false;
    std::stringstream ss(line.substr(a + 1, b - a - 1));
    std::string tok;
    std::vector<double> v;
    while (std::getline(ss, tok, ',')) v.push_back(strtod(tok.c_str(), nullptr));
    if ((int)v.size() < n) return false;
    std::copy_n(v.begin(), n, out);
    return true;
}

void lerp3(const double* a, const double* b, double t, double* o) { for (int i = 0; i < 3; ++i) o[i] = (1 - t) * a[i] + t * b[i]; }

void slerp(const double* a, const double* b_in, double t, double* o)
{
    double b[4] = {b_in[0], b_in[1], b_in[2], b_in[3]};
    double d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
    if (d < 0) { d = -d; for (double& v : b) v = -v; }        // shorter arc
    double wa, wb;
    if (d > 1.0 - 1e-12) { wa = 1 - t; wb = t; }              // nearly parallel: linear
    else {
        const double th = std::acos(d), s = std::sin(th);
        wa = std::sin((1 - t) * th) / s; wb = std::sin(t * th) / s;
    }
    for (int i = 0; i < 4; ++i) o[i] = wa * a[i] + wb * b[i];
}

}  // namespace

// ---------------------------------------------------------------- PNG
bool png_decode_gray8(const uint8_t* file, size_t n, uint8_t* dst, size_t dst_bytes, int& width, int& height, std::string* err)
{
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (n < 8 + 25 || memcmp(file, sig, 8) != 0) return set_err(err, "not a PNG file");
    size_t pos = 8;
    bool have_hdr = false, done = false;
    std::vector<uint8_t> idat;
    while (!done && pos + 12 <= n) {
        const uint32_t len = be32(file + pos);
        const uint8_t* type = file + pos + 4;
        if (pos + 12 + (size_t)len > n) return set_err(err, "truncated PNG chunk");
        const uint32_t crc = be32(file + pos + 8 + len);
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), type, 4 + len) != crc) return set_err(err, "PNG chunk CRC mismatch");
        const uint8_t* body = file + pos + 8;
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) return set_err(err, "bad IHDR");
            width = (int)be32(body); height = (int)be32(body + 4);
            if (width <= 0 || height <= 0) return set_err(err, "bad PNG size");
            if (body[8] != 8 || body[9] != 0 || body[12] != 0) return set_err(err, "unsupported PNG (need 8-bit greyscale, non-interlaced)");
            have_hdr = true;
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!memcmp(type, "IEND", 4)) {
            done = true;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_hdr || !done) return set_err(err, "PNG without IHDR/IEND");
    if (!dst) return true;
    const size_t w = (size_t)width, h = (size_t)height;
    if (dst_bytes < w * h) return set_err(err, "destination too small");
    std::vector<uint8_t> raw((w + 1) * h);
    uLongf got = (uLongf)raw.size();
    if (uncompress(raw.data(), &got, idat.data(), (uLong)idat.size()) != Z_OK || got != raw.size()) return set_err(err, "bad zlib stream in PNG");
    // un-filter, bytes per pixel = 1: a = left, b = up, c = up-left
    for (size_t y = 0; y < h; ++y) {
        const uint8_t* src = raw.data() + y * (w + 1);
        uint8_t* row = dst + y * w;
        const uint8_t* up = y ? row - w : nullptr;
        switch (src[0]) {
        case 0: memcpy(row, src + 1, w); break;
        case 1: row[0] = src[1]; for (size_t x = 1; x < w; ++x) row[x] = (uint8_t)(src[x + 1] + row[x - 1]); break;
        case 2: for (size_t x = 0; x < w; ++x) row[x] = (uint8_t)(src[x + 1] + (up ? up[x] : 0)); break;
        case 3: for (size_t x = 0; x < w; ++x) row[x] = (uint8_t)(src[x + 1] + (((x ? row[x - 1] : 0) + (up ? up[x] : 0)) >> 1)); break;
        case 4:
            for (size_t x = 0; x < w; ++x) {
                const int a = x ? row[x - 1] : 0, b = up ? up[x] : 0, c = (x && up) ? up[x - 1] : 0;
                const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                row[x] = (uint8_t)(src[x + 1] + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)));
            }
            break;
        default: return set_err(err, "bad PNG filter type");
        }
    }
    return true;
}

// ---------------------------------------------------------------- reader
EurocReader::EurocReader(std::string dataset_path) : root_(std::move(dataset_path)) {}

bool EurocReader::load()
{
    std::string mav = root_ + "/mav0";
    if (!file_exists(mav + "/cam0/data.csv")) mav = root_;
    stamps_.clear(); files_.clear(); imu_.clear(); gt_.clear();
    reset();

    std::vector<CsvRow> rows;
    if (read_csv(mav + "/cam0/data.csv", 2, 1, rows)) {
        std::stable_sort(rows.begin(), rows.end(), [](const CsvRow& a, const CsvRow& b) { return a.num[0] < b.num[0]; });
        for (const auto& r : rows) { stamps_.push_back(r.num[0] * 1e-9); files_.push_back(mav + "/cam0/data/" + r.text); }
    }
    rows.clear();
    bool imu_ok = read_csv(mav + "/imu0/data.csv", 7, -1, rows);
    std::stable_sort(rows.begin(), rows.end(), [](const CsvRow& a, const CsvRow& b) { return a.num[0] < b.num[0]; });
    for (const auto& r : rows) {
        EurocImu m;                                            // timestamp, gyro xyz [rad/s], accel xyz [m/s^2]
        m.timestamp = r.num[0] * 1e-9;
        for (int i = 0; i < 3; ++i) { m.gyro[i] = r.num[1 + i]; m.accel[i] = r.num[4 + i]; }
        imu_.push_back(m);
    }
    imu_ok = imu_ok && !imu_.empty();
    rows.clear();
    if (!read_csv(mav + "/state_groundtruth_estimate0/data.csv", 17, -1, rows) || rows.empty()) {
        rows.clear();
        read_csv(mav + "/leica0/data.csv", 17, -1, rows);
    }
    std::stable_sort(rows.begin(), rows.end(), [](const CsvRow& a, const CsvRow& b) { return a.num[0] < b.num[0]; });
    for (const auto& r : rows) {
        EurocGroundTruth g;                                    // t, p(3), q(w x y z), v(3), bias gyro(3), bias accel(3)
        g.timestamp = r.num[0] * 1e-9;
        for (int i = 0; i < 3; ++i) { g.position[i] = r.num[1 + i]; g.velocity[i] = r.num[8 + i]; g.bias_gyro[i] = r.num[11 + i]; g.bias_accel[i] = r.num[14 + i]; }
        for (int i = 0; i < 4; ++i) g.orientation[i] = r.num[4 + i];
        gt_.push_back(g);
    }
    std::ifstream yaml(mav + "/cam0/sensor.yaml");
    std::string line;
    while (yaml.is_open() && std::getline(yaml, line)) {
        yaml_list(line, "intrinsics:", K_, 4);
        yaml_list(line, "distortion_coefficients:", dist_, 4);
    }
    return imu_ok && !stamps_.empty();                         // as the reference: images and IMU are required
}

bool EurocReader::readImage(size_t i, uint8_t* dst, size_t dst_bytes, int& width, int& height, std::string* err) const
{
    if (i >= files_.size()) return set_err(err, "frame index out of range");
    std::vector<uint8_t> file;
    if (!read_file(files_[i], file)) return set_err(err, "cannot read " + files_[i]);
    return png_decode_gray8(file.data(), file.size(), dst, dst_bytes, width, height, err);
}

bool EurocReader::imageSize(size_t i, int& width, int& height) const { return readImage(i, nullptr, 0, width, height); }

void EurocReader::imuBetween(double prev, double now, std::vector<EurocImu>& out)
{
    out.clear();
    while (imu_cursor_ < imu_.size() && imu_[imu_cursor_].timestamp <= now) {
        if (imu_[imu_cursor_].timestamp > prev) out.push_back(imu_[imu_cursor_]);
        ++imu_cursor_;
    }
}

bool EurocReader::getNext(std::vector<uint8_t>& gray, int& width, int& height, std::vector<EurocImu>& imu, double& timestamp)
{
    while (cursor_ < stamps_.size()) {
        const size_t i = cursor_++;
        int w = 0, h = 0;
        if (!imageSize(i, w, h)) continue;                     // unreadable frame: skip, like the reference
        gray.resize((size_t)w * h);
        if (!readImage(i, gray.data(), gray.size(), w, h)) continue;
        width = w; height = h;
        timestamp = stamps_[i];
        imuBetween(i ? stamps_[i - 1] : 0.0, timestamp, imu);
        return true;
    }
    return false;
}

bool EurocReader::getGroundTruth(double t, EurocGroundTruth& out) const
{
    if (gt_.empty()) return false;
    auto it = std::lower_bound(gt_.begin(), gt_.end(), t, [](const EurocGroundTruth& g, double v) { return g.timestamp < v; });
    if (it == gt_.end()) { out = gt_.back(); return true; }
    if (it == gt_.begin()) { out = gt_.front(); return true; }
    const EurocGroundTruth &a = *(it - 1), &b = *it;
    const double u = (t - a.timestamp) / (b.timestamp - a.timestamp);
    out.timestamp = t;
    lerp3(a.position, b.position, u, out.position);
    lerp3(a.velocity, b.velocity, u, out.velocity);
    lerp3(a.bias_gyro, b.bias_gyro, u, out.bias_gyro);
    lerp3(a.bias_accel, b.bias_accel, u, out.bias_accel);
    slerp(a.orientation, b.orientation, u, out.orientation);
    return true;
}

// ---------------------------------------------------------------- replay
namespace {

struct Pinned {
    void* p = nullptr;
    ~Pinned() { if (p) aslam_host_free(p); }
    bool alloc(size_t n) { return aslam_host_alloc(n, &p) == ASLAM_OK; }
};

// decode frames [lo, hi) into buf (frame stride w*h) on `threads` host threads; returns summed decode seconds or < 0
double decode_batch(const EurocReader& rd, size_t lo, size_t hi, uint8_t* buf, int w, int h, int threads, std::string* err)
{
    std::atomic<size_t> next{lo};
    std::atomic<bool> ok{true};
    std::vector<double> secs((size_t)threads, 0.0);
    std::vector<std::string> errs((size_t)threads);
    auto work = [&](int t) {
        const auto t0 = std::chrono::steady_clock::now();
        for (size_t i = next++; i < hi && ok; i = next++) {
            int fw = 0, fh = 0;
            if (!rd.readImage(i, buf + (i - lo) * (size_t)w * h, (size_t)w * h, fw, fh, &errs[(size_t)t]) || fw != w || fh != h) {
                if (errs[(size_t)t].empty()) errs[(size_t)t] = "frame size differs within the sequence: " + rd.filename(i);
                ok = false;
            }
        }
        secs[(size_t)t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    if (!ok) { for (const auto& e : errs) if (!e.empty()) { set_err(err, e); break; } return -1.0; }
    double s = 0;
    for (double v : secs) s += v;
    return s;
}

}  // namespace

bool euroc_replay(const EurocReader& reader, aslam_ctx* ctx, int nfeatures, int batch, int threads, size_t first, size_t count,
                  const ReplayFrameFn& on_frame, ReplayStats& stats, std::string* err)
{
    stats = ReplayStats();
    if (first >= reader.numImages() || batch <= 0) return set_err(err, "nothing to replay");
    count = std::min(count, reader.numImages() - first);
    threads = std::max(1, threads);
    int w = 0, h = 0;
    if (!reader.imageSize(first, w, h)) return set_err(err, "cannot read the first frame: " + reader.filename(first));
    const auto t_start = std::chrono::steady_clock::now();

    aslam_orb_params prm;
    aslam_orb_default_params(&prm);
    prm.nfeatures = nfeatures;
    aslam_orb* orb = nullptr;
    if (aslam_orb_create(ctx, &prm, w, h, batch, &orb) != ASLAM_OK) return set_err(err, aslam_last_error());
    const size_t cap = (size_t)aslam_orb_capacity(orb), fs = (size_t)w * h;
    Pinned img[2], kps, desc, cnt, mat, mcnt;
    bool ok = img[0].alloc(fs * batch) && img[1].alloc(fs * batch) && kps.alloc(sizeof(aslam_keypoint) * cap * batch) &&
              desc.alloc(32 * cap * batch) && cnt.alloc(sizeof(int32_t) * batch) && mat.alloc(sizeof(aslam_match) * cap * batch) &&
              mcnt.alloc(sizeof(int32_t) * batch);
    if (!ok) { aslam_orb_destroy(orb); return set_err(err, "pinned allocation failed"); }

    const size_t nb = (count + (size_t)batch - 1) / (size_t)batch;
    auto range = [&](size_t b, size_t& lo, size_t& hi) { lo = first + b * (size_t)batch; hi = std::min(lo + (size_t)batch, first + count); };
    size_t lo, hi;
    range(0, lo, hi);
    double ds = decode_batch(reader, lo, hi, (uint8_t*)img[0].p, w, h, threads, err);
    ok = ds >= 0;
    for (size_t b = 0; ok && b < nb; ++b) {
        stats.decode_s += ds;
        range(b, lo, hi);
        const int n = (int)(hi - lo);
        if (aslam_orb_track_batch_async(orb, (const uint8_t*)img[b & 1].p, n, fs, w, 0.75f, 0) != ASLAM_OK) { ok = set_err(err, aslam_last_error()); break; }
        if (b + 1 < nb) {                                      // decode the next batch while the GPU runs this one
            size_t nlo, nhi;
            range(b + 1, nlo, nhi);
            ds = decode_batch(reader, nlo, nhi, (uint8_t*)img[(b + 1) & 1].p, w, h, threads, err);
            if (ds < 0) { ok = false; aslam_ctx_sync(ctx); break; }
        }
        const int rc = aslam_orb_track_sync(orb, (aslam_keypoint*)kps.p, (uint8_t*)desc.p, (int32_t*)cnt.p, (aslam_match*)mat.p, (int32_t*)mcnt.p);
        if (rc != ASLAM_OK && rc != ASLAM_E_OVERFLOW) { ok = set_err(err, aslam_last_error()); break; }
        for (int i = 0; i < n; ++i) {
            const int nk = ((int32_t*)cnt.p)[i], nm = ((int32_t*)mcnt.p)[i];
            stats.keypoints += (size_t)nk; stats.matches += (size_t)nm;
            if (on_frame) on_frame(lo + (size_t)i, reader.timestamp(lo + (size_t)i), (aslam_keypoint*)kps.p + (size_t)i * cap,
                                   (uint8_t*)desc.p + (size_t)i * cap * 32, nk, (aslam_match*)mat.p + (size_t)i * cap, nm);
        }
        stats.frames += (size_t)n;
        ++stats.batches;
    }
    aslam_orb_destroy(orb);
    stats.total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    return ok;
}

}  // namespace aria::adapters::b200
