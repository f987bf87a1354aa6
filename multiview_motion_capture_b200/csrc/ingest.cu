// Ingest: OpenPose JSON -> BODY_25 arrays -> COCO-17 poses packed for the clip pipeline.
//
// Reference: motion_capture.py:974-1005 (parse_openpose_kps / extract_frame_data_from_openpose: json.load of every
// `*_keypoints.json`, `people[].pose_keypoints_2d` reshaped (25, 3)) and pose_def.py:262-270
// (conversion_openpose_25_to_coco: a 17-joint gather). The reference does this per frame in Python and pickles the result;
// here the JSON text is scanned by a native parser (one thread per file, correctly rounded strtod: the same doubles
// json.load produces) straight into the packed layout the device wants, and the joint gather + padding runs on the device.
#include "mvmc_common.cuh"

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

namespace mvmc {

// COCO slot <- BODY_25 slot (pose_def.py:262-270)
__constant__ int c_b25_to_coco[MVMC_N_COCO] = {0, 16, 15, 18, 17, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11};

// kps25 [B,C,Pin,25,3] -> kps [B,C,Pmax,17,3] (zero beyond n_people), n_pose = min(n_people, Pmax)
__global__ void __launch_bounds__(256)
    k_body25_to_coco(const double* __restrict__ kps25, const int* __restrict__ n_people, int BC, int Pin, int Pmax,
                     double* __restrict__ kps, int* __restrict__ n_pose) {
    const int per = Pmax * MVMC_N_COCO * 3;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)BC * per; e += (long long)gridDim.x * blockDim.x) {
        const int bc = (int)(e / per), r = (int)(e % per);
        const int p = r / (MVMC_N_COCO * 3), j = (r / 3) % MVMC_N_COCO, c = r % 3;
        const int np = min(n_people[bc], min(Pin, Pmax));
        double v = 0.0;
        if (p < np) v = kps25[(((size_t)bc * Pin + p) * 25 + c_b25_to_coco[j]) * 3 + c];
        kps[e] = v;
        if (r == 0) n_pose[bc] = np;
    }
}

// ---- host side: a scanner for OpenPose's JSON (objects, arrays, strings, numbers; no dependency) ----
struct Scan {
    const char* p;
    const char* end;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool lit(char c) {
        ws();
        if (p < end && *p == c) {
            p++;
            return true;
        }
        return false;
    }
    bool string(std::string* out) {
        ws();
        if (p >= end || *p != '"') return false;
        p++;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) p++;
            if (out) out->push_back(*p);
            p++;
        }
        if (p >= end) return false;
        p++;
        return true;
    }
    bool skip_value() {   // any JSON value
        ws();
        if (p >= end) return false;
        if (*p == '"') return string(nullptr);
        if (*p == '{' || *p == '[') {
            const char open = *p, close = open == '{' ? '}' : ']';
            p++;
            ws();
            if (lit(close)) return true;
            for (;;) {
                if (open == '{') {
                    if (!string(nullptr) || !lit(':')) return false;
                }
                if (!skip_value()) return false;
                if (lit(',')) continue;
                return lit(close);
            }
        }
        while (p < end && *p != ',' && *p != '}' && *p != ']' && *p != ' ' && *p != '\n' && *p != '\r' && *p != '\t') p++;
        return true;
    }
    // array of numbers -> out[0..cap); returns the count or -1
    int numbers(double* out, int cap) {
        if (!lit('[')) return -1;
        int n = 0;
        ws();
        if (lit(']')) return 0;
        for (;;) {
            ws();
            char* q = nullptr;
            const double v = strtod(p, &q);
            if (q == p || q > end) return -1;
            if (n < cap) out[n] = v;
            n++;
            p = q;
            if (lit(',')) continue;
            return lit(']') ? n : -1;
        }
    }
};

// text of one OpenPose file -> out [max_people][25][3], *n_people = len(people) (may exceed max_people: extra ones dropped)
static int parse_openpose(const char* text, size_t len, int max_people, double* out, int* n_people) {
    Scan s{text, text + len};
    *n_people = 0;
    if (!s.lit('{')) return MVMC_ERR_INVALID;
    if (s.lit('}')) return MVMC_OK;
    for (;;) {
        std::string key;
        if (!s.string(&key) || !s.lit(':')) return MVMC_ERR_INVALID;
        if (key == "people") {
            if (!s.lit('[')) return MVMC_ERR_INVALID;
            int np = 0;
            s.ws();
            if (!s.lit(']')) {
                for (;;) {   // one person object
                    if (!s.lit('{')) return MVMC_ERR_INVALID;
                    bool seen = false;
                    if (!s.lit('}')) {
                        for (;;) {
                            std::string k2;
                            if (!s.string(&k2) || !s.lit(':')) return MVMC_ERR_INVALID;
                            if (k2 == "pose_keypoints_2d") {
                                double tmp[75];
                                const int cnt = s.numbers(tmp, 75);
                                if (cnt != 75) return MVMC_ERR_INVALID;   // BODY_25 (the reference reshapes (-1, 3) and gathers slot 18)
                                if (np < max_people) memcpy(out + (size_t)np * 75, tmp, sizeof(tmp));
                                seen = true;
                            } else if (!s.skip_value()) {
                                return MVMC_ERR_INVALID;
                            }
                            if (s.lit(',')) continue;
                            if (!s.lit('}')) return MVMC_ERR_INVALID;
                            break;
                        }
                    }
                    if (!seen) return MVMC_ERR_INVALID;
                    np++;
                    if (s.lit(',')) continue;
                    if (!s.lit(']')) return MVMC_ERR_INVALID;
                    break;
                }
            }
            *n_people = np;
        } else if (!s.skip_value()) {
            return MVMC_ERR_INVALID;
        }
        if (s.lit(',')) continue;
        return s.lit('}') ? MVMC_OK : MVMC_ERR_INVALID;
    }
}

}  // namespace mvmc

using namespace mvmc;

extern "C" int mvmc_parse_openpose_host(const char* text, size_t len, int max_people, double* out, int* n_people) {
    if (!text || !out || !n_people || max_people <= 0) return MVMC_ERR_INVALID;
    memset(out, 0, (size_t)max_people * 75 * sizeof(double));
    return parse_openpose(text, len, max_people, out, n_people);
}

extern "C" int mvmc_parse_openpose_files_host(const char* const* paths, int n_files, int max_people, double* out, int* n_people,
                                              int n_threads) {
    if (!paths || !out || !n_people || n_files <= 0 || max_people <= 0) return MVMC_ERR_INVALID;
    std::atomic<int> next(0), bad(0);
    auto work = [&]() {
        std::vector<char> buf;
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n_files) break;
            double* o = out + (size_t)i * max_people * 75;
            memset(o, 0, (size_t)max_people * 75 * sizeof(double));
            n_people[i] = 0;
            FILE* f = fopen(paths[i], "rb");
            if (!f) {
                bad.store(1);
                continue;
            }
            fseek(f, 0, SEEK_END);
            const long sz = ftell(f);
            fseek(f, 0, SEEK_SET);
            buf.resize(sz > 0 ? (size_t)sz + 1 : 1);
            const size_t got = sz > 0 ? fread(buf.data(), 1, (size_t)sz, f) : 0;
            fclose(f);
            buf[got] = 0;   // strtod needs a terminator
            if (parse_openpose(buf.data(), got, max_people, o, &n_people[i]) != MVMC_OK) bad.store(1);
        }
    };
    const int nt = n_threads < 1 ? 1 : (n_threads > n_files ? n_files : n_threads);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return bad.load() ? MVMC_ERR_INVALID : MVMC_OK;
}

extern "C" int mvmc_ingest_body25(const double* kps25, const int* n_people, int B, int C, int Pin, int Pmax, double* kps,
                                  int* n_pose, void* stream) {
    if (!kps25 || !n_people || !kps || !n_pose) return MVMC_ERR_INVALID;
    if (B <= 0 || C <= 0 || C > MVMC_MAX_VIEWS || Pin <= 0 || Pmax <= 0 || Pmax > MVMC_MAX_POSES) return MVMC_ERR_INVALID;
    const long long total = (long long)B * C * Pmax * MVMC_N_COCO * 3;
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    MVMC_LAUNCH(k_body25_to_coco, dim3(grid), dim3(256), 0, stream, kps25, n_people, B * C, Pin, Pmax, kps, n_pose);
    MVMC_CHECK_LAUNCH("k_body25_to_coco");
    return MVMC_OK;
}
