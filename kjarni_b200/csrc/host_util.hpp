// Host-side support for libkjarni_cuda: thread-local error state, a small JSON
// reader (config.json, safetensors header, segment.json) and an mmap'ed
// safetensors container reader.  No third-party dependencies.
#pragma once
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kjarni_cuda.h"

namespace kj {

// ------------------------------------------------------------------ errors
struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};
void set_last_error(const std::string& msg);  // thread-local (capi.cu)

#define KJ_CUDA(expr)                                                                                            \
    do {                                                                                                         \
        cudaError_t e__ = (expr);                                                                                \
        if (e__ != cudaSuccess)                                                                                  \
            throw ::kj::Error(KJC_INFERENCE_FAILED, std::string("CUDA error: ") + cudaGetErrorString(e__) + " at " + \
                                                        __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")"); \
    } while (0)

// -------------------------------------------------------------------- JSON
struct Json {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;  // insertion order kept

    const Json* get(const std::string& k) const {
        if (type != Obj) return nullptr;
        for (auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
    bool has(const std::string& k) const {
        const Json* j = get(k);
        return j && j->type != Null;
    }
    double number(const std::string& k, double dflt) const {
        const Json* j = get(k);
        return (j && j->type == Num) ? j->num : dflt;
    }
    std::string string(const std::string& k, const std::string& dflt) const {
        const Json* j = get(k);
        return (j && j->type == Str) ? j->str : dflt;
    }
};

class JsonParser {
  public:
    JsonParser(const char* s, size_t n) : p_(s), end_(s + n) {}
    Json parse() {
        Json j = value();
        ws();
        if (p_ != end_) fail("trailing characters");
        return j;
    }

  private:
    const char* p_;
    const char* end_;
    [[noreturn]] void fail(const char* m) { throw Error(KJC_INVALID_CONFIG, std::string("JSON parse error: ") + m); }
    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
    }
    Json value() {
        ws();
        if (p_ >= end_) fail("unexpected end");
        Json j;
        const char c = *p_;
        if (c == '{') {
            j.type = Json::Obj;
            ++p_;
            ws();
            if (p_ < end_ && *p_ == '}') { ++p_; return j; }
            while (true) {
                ws();
                if (p_ >= end_ || *p_ != '"') fail("expected key");
                std::string k = str();
                ws();
                if (p_ >= end_ || *p_ != ':') fail("expected ':'");
                ++p_;
                Json v = value();
                j.obj.emplace_back(std::move(k), std::move(v));
                ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == '}') { ++p_; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            j.type = Json::Arr;
            ++p_;
            ws();
            if (p_ < end_ && *p_ == ']') { ++p_; return j; }
            while (true) {
                j.arr.push_back(value());
                ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == ']') { ++p_; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            j.type = Json::Str;
            j.str = str();
        } else if (c == 't' && end_ - p_ >= 4 && !strncmp(p_, "true", 4)) {
            j.type = Json::Bool; j.b = true; p_ += 4;
        } else if (c == 'f' && end_ - p_ >= 5 && !strncmp(p_, "false", 5)) {
            j.type = Json::Bool; j.b = false; p_ += 5;
        } else if (c == 'n' && end_ - p_ >= 4 && !strncmp(p_, "null", 4)) {
            j.type = Json::Null; p_ += 4;
        } else if (c == 'N' && end_ - p_ >= 3 && !strncmp(p_, "NaN", 3)) {  // python json.dump leniency
            j.type = Json::Num; j.num = NAN; p_ += 3;
        } else {
            char* e = nullptr;
            std::string tmp(p_, std::min<size_t>(end_ - p_, 64));
            j.num = strtod(tmp.c_str(), &e);
            if (e == tmp.c_str()) fail("bad value");
            j.type = Json::Num;
            p_ += (e - tmp.c_str());
        }
        return j;
    }
    static void utf8(std::string& o, unsigned cp) {
        if (cp < 0x80) o += char(cp);
        else if (cp < 0x800) { o += char(0xC0 | (cp >> 6)); o += char(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { o += char(0xE0 | (cp >> 12)); o += char(0x80 | ((cp >> 6) & 0x3F)); o += char(0x80 | (cp & 0x3F)); }
        else { o += char(0xF0 | (cp >> 18)); o += char(0x80 | ((cp >> 12) & 0x3F)); o += char(0x80 | ((cp >> 6) & 0x3F)); o += char(0x80 | (cp & 0x3F)); }
    }
    unsigned hex4() {
        if (end_ - p_ < 4) fail("bad \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else fail("bad hex digit");
        }
        return v;
    }
    std::string str() {
        ++p_;  // opening quote
        std::string o;
        while (true) {
            if (p_ >= end_) fail("unterminated string");
            const char c = *p_++;
            if (c == '"') break;
            if (c != '\\') { o += c; continue; }
            if (p_ >= end_) fail("bad escape");
            const char e = *p_++;
            switch (e) {
                case '"': o += '"'; break;
                case '\\': o += '\\'; break;
                case '/': o += '/'; break;
                case 'b': o += '\b'; break;
                case 'f': o += '\f'; break;
                case 'n': o += '\n'; break;
                case 'r': o += '\r'; break;
                case 't': o += '\t'; break;
                case 'u': {
                    unsigned cp = hex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                        p_ += 2;
                        const unsigned lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    utf8(o, cp);
                    break;
                }
                default: fail("unknown escape");
            }
        }
        return o;
    }
};

inline std::string read_text_file(const std::string& path, int missing_status) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Error(missing_status, "cannot open " + path);
    std::string s;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
    fclose(f);
    return s;
}

// ------------------------------------------------------------ safetensors
// Standard container: u64 LE header length, JSON header {name: {dtype, shape, data_offsets}}, raw LE data
// (reference: kjarni-transformers/src/weights/safetensors_loader.rs:131-176; dtypes accepted:
// tensor/dtype.rs:28-39).  The file is mmap'ed read-only, as the reference does.
struct StTensor {
    std::string dtype;
    std::vector<int64_t> shape;
    const uint8_t* data = nullptr;
    size_t nbytes = 0;
    int64_t numel() const {
        int64_t n = 1;
        for (auto d : shape) n *= d;
        return n;
    }
};

class SafeTensors {
  public:
    explicit SafeTensors(const std::string& path) {
        fd_ = open(path.c_str(), O_RDONLY);
        if (fd_ < 0) throw Error(KJC_MODEL_NOT_FOUND, "cannot open " + path);
        struct stat st;
        if (fstat(fd_, &st) != 0 || st.st_size < 8) { close(fd_); throw Error(KJC_LOAD_FAILED, "bad safetensors file " + path); }
        size_ = static_cast<size_t>(st.st_size);
        map_ = static_cast<uint8_t*>(mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0));
        if (map_ == MAP_FAILED) { close(fd_); throw Error(KJC_LOAD_FAILED, "mmap failed for " + path); }
        uint64_t hlen;
        memcpy(&hlen, map_, 8);
        if (hlen > size_ - 8) { cleanup(); throw Error(KJC_LOAD_FAILED, "safetensors header length out of range: " + path); }
        try {
            Json h = JsonParser(reinterpret_cast<const char*>(map_ + 8), hlen).parse();
            if (h.type != Json::Obj) throw Error(KJC_LOAD_FAILED, "safetensors header is not an object");
            const uint8_t* base = map_ + 8 + hlen;
            const size_t avail = size_ - 8 - hlen;
            for (auto& kv : h.obj) {
                if (kv.first == "__metadata__") continue;
                StTensor t;
                t.dtype = kv.second.string("dtype", "");
                const Json* sh = kv.second.get("shape");
                const Json* off = kv.second.get("data_offsets");
                if (!sh || sh->type != Json::Arr || !off || off->type != Json::Arr || off->arr.size() != 2)
                    throw Error(KJC_LOAD_FAILED, "malformed safetensors entry " + kv.first);
                for (auto& d : sh->arr) t.shape.push_back(static_cast<int64_t>(d.num));
                const size_t a = static_cast<size_t>(off->arr[0].num), b = static_cast<size_t>(off->arr[1].num);
                if (a > b || b > avail) throw Error(KJC_LOAD_FAILED, "safetensors offsets out of range for " + kv.first);
                t.data = base + a;
                t.nbytes = b - a;
                tensors_[kv.first] = std::move(t);
            }
        } catch (const Error& e) {
            cleanup();
            throw Error(KJC_LOAD_FAILED, std::string(e.what()));
        }
    }
    ~SafeTensors() { cleanup(); }
    SafeTensors(const SafeTensors&) = delete;
    SafeTensors& operator=(const SafeTensors&) = delete;

    bool contains(const std::string& name) const { return tensors_.count(name) != 0; }
    const StTensor& at(const std::string& name) const {
        auto it = tensors_.find(name);
        if (it == tensors_.end()) throw Error(KJC_LOAD_FAILED, "missing tensor '" + name + "'");
        return it->second;
    }
    // Tensor as fp32 (F32 copied; F16 / BF16 up-cast, linear_layer/builder.rs:108-138).
    std::vector<float> as_f32(const std::string& name) const {
        const StTensor& t = at(name);
        const int64_t n = t.numel();
        std::vector<float> out(static_cast<size_t>(n));
        if (t.dtype == "F32") {
            if (t.nbytes != static_cast<size_t>(n) * 4) throw Error(KJC_LOAD_FAILED, "size mismatch for " + name);
            memcpy(out.data(), t.data, t.nbytes);
        } else if (t.dtype == "BF16") {
            if (t.nbytes != static_cast<size_t>(n) * 2) throw Error(KJC_LOAD_FAILED, "size mismatch for " + name);
            for (int64_t i = 0; i < n; ++i) {
                uint16_t h;
                memcpy(&h, t.data + 2 * i, 2);
                const uint32_t u = static_cast<uint32_t>(h) << 16;
                memcpy(&out[i], &u, 4);
            }
        } else if (t.dtype == "F16") {
            if (t.nbytes != static_cast<size_t>(n) * 2) throw Error(KJC_LOAD_FAILED, "size mismatch for " + name);
            for (int64_t i = 0; i < n; ++i) {
                uint16_t h;
                memcpy(&h, t.data + 2 * i, 2);
                const uint32_t sign = (h & 0x8000u) << 16;
                uint32_t exp = (h >> 10) & 0x1F, man = h & 0x3FF, u;
                if (exp == 0) {
                    if (man == 0) u = sign;
                    else {
                        int e = -1;
                        do { ++e; man <<= 1; } while (!(man & 0x400));
                        u = sign | ((127 - 15 - e) << 23) | ((man & 0x3FF) << 13);
                    }
                } else if (exp == 31) u = sign | 0x7F800000u | (man << 13);
                else u = sign | ((exp + 112) << 23) | (man << 13);
                memcpy(&out[i], &u, 4);
            }
        } else {
            throw Error(KJC_LOAD_FAILED, "unsupported dtype " + t.dtype + " for " + name);
        }
        return out;
    }

  private:
    void cleanup() {
        if (map_ && map_ != MAP_FAILED) munmap(map_, size_);
        map_ = nullptr;
        if (fd_ >= 0) close(fd_);
        fd_ = -1;
    }
    int fd_ = -1;
    uint8_t* map_ = nullptr;
    size_t size_ = 0;
    std::map<std::string, StTensor> tensors_;
};

}  // namespace kj
