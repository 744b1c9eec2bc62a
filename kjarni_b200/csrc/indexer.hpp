// Host side of the indexing pipeline around the GPU embedder: text splitting, file discovery and the on-disk
// segment / index writer, restated from
//   TextSplitter::{split, get_overlap_suffix, split_large_text}      kjarni-rag/src/splitter.rs:59-163
//   DocumentLoader::load_file, TEXT_EXTENSIONS, ChunkMetadata        kjarni-rag/src/loader.rs:6-93, kjarni-search/src/types.rs:56-78
//   Indexer::collect_files / is_supported_file                       kjarni/src/indexer/model.rs:727-800
//   SegmentBuilder::{new, add, flush}, IndexWriter::{open, open_existing, add, commit}
//                                                                    kjarni-rag/src/segment.rs:45-193, index_writer.rs:18-170
// What differs on purpose: embeddings are NOT passed through a Vec<Vec<f32>>: the writer hands out the byte offset of a row inside the
// segment's vectors file and the embedder's pinned output buffer is pwrite()n there directly; directory walks are sorted by path
// (WalkDir yields readdir order, which is unspecified) so an index build is reproducible.
#pragma once
#include <dirent.h>
#include <fcntl.h>
#include <fnmatch.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <ctime>
#include <map>
#include <string>
#include <vector>

#include "bm25.hpp"

namespace kj {

// ----------------------------------------------------------------- TextSplitter
struct TextSplitter {
    size_t chunk_size = 1000, chunk_overlap = 200;  // lengths in BYTES for the section logic, in CHARS for the overlap / hard split
    std::string separator = "\n\n";

    static std::string chars_to_string(const std::vector<uint32_t>& c, size_t a, size_t b) {
        std::string s;
        for (size_t i = a; i < b; ++i) uni::encode_append(c[i], s);
        return s;
    }
    std::string overlap_suffix(const std::string& text) const {  // get_overlap_suffix, splitter.rs:113-122
        const std::vector<uint32_t> c = uni::decode(text);
        if (c.size() <= chunk_overlap) return text;
        return chars_to_string(c, c.size() - chunk_overlap, c.size());
    }
    void split_large(const std::string& text, std::vector<std::string>& out) const {  // split_large_text, splitter.rs:124-163
        const std::vector<uint32_t> c = uni::decode(text);
        if (c.empty()) return;
        size_t start = 0;
        while (start < c.size()) {
            const size_t end = std::min(start + chunk_size, c.size());
            out.push_back(chars_to_string(c, start, end));
            if (end >= c.size()) break;
            const size_t step = (chunk_overlap > 0 && chunk_overlap < chunk_size) ? chunk_size - chunk_overlap : chunk_size;
            start = step > 0 ? start + step : start + 1;
        }
    }
    std::vector<std::string> split(const std::string& text) const {  // split, splitter.rs:59-111
        std::vector<std::string> chunks;
        if (text.empty()) return chunks;
        std::string cur;
        size_t pos = 0;
        for (;;) {
            const size_t nx = separator.empty() ? std::string::npos : text.find(separator, pos);
            const std::string section = text.substr(pos, nx == std::string::npos ? std::string::npos : nx - pos);
            if (!section.empty()) {
                if (section.size() > chunk_size) {
                    if (!cur.empty()) {
                        chunks.push_back(cur);
                        cur.clear();
                    }
                    split_large(section, chunks);
                } else {
                    const size_t would = cur.empty() ? section.size() : cur.size() + separator.size() + section.size();
                    if (would > chunk_size && !cur.empty()) {
                        chunks.push_back(cur);
                        if (chunk_overlap > 0) cur = overlap_suffix(cur);
                        else cur.clear();
                    }
                    if (!cur.empty()) cur += separator;
                    cur += section;
                }
            }
            if (nx == std::string::npos) break;
            pos = nx + separator.size();
        }
        if (!cur.empty()) chunks.push_back(cur);
        return chunks;
    }
};

// ----------------------------------------------------------------- file discovery
struct LoaderOptions {
    std::vector<std::string> extensions;        // lower-case, without the dot; empty = the reference's TEXT_EXTENSIONS
    std::vector<std::string> exclude_patterns;  // globs matched against the whole path
    bool recursive = true, include_hidden = false;
    size_t max_file_size = 10u * 1024 * 1024;   // 0 = no limit
};

inline bool supported_extension(const std::string& path, const LoaderOptions& o) {
    static const char* kText[] = {"txt", "md", "markdown", "rst", "org", "json", "yaml", "yml", "toml", "xml", "csv", "html", "htm", "css", "rs", "py",
                                  "js", "ts", "go", "java", "c", "cpp", "h", "hpp", "cs", "rb", "sh", "bash", "zsh", "fish", "ps1", "sql", "r", "scala",
                                  "kt", "swift", "m", "mm", "lua", "pl", "php", "ex", "exs", "clj", "hs"};
    const size_t slash = path.find_last_of('/');
    const std::string name = slash == std::string::npos ? path : path.substr(slash + 1);
    const size_t dot = name.find_last_of('.');
    if (dot == std::string::npos || dot == 0) return false;  // Path::extension(): none for "file" and for ".hidden"
    std::string ext = name.substr(dot + 1);
    for (char& c : ext) c = static_cast<char>(tolower(static_cast<unsigned char>(c)));
    if (o.extensions.empty()) {
        for (const char* e : kText)
            if (ext == e) return true;
        return false;
    }
    return std::find(o.extensions.begin(), o.extensions.end(), ext) != o.extensions.end();
}

inline void walk_dir(const std::string& dir, bool recursive, std::vector<std::string>& files) {
    DIR* d = opendir(dir.c_str());
    if (!d) return;
    std::vector<std::string> names;
    while (dirent* e = readdir(d)) {
        const std::string n = e->d_name;
        if (n != "." && n != "..") names.push_back(n);
    }
    closedir(d);
    std::sort(names.begin(), names.end());
    for (const std::string& n : names) {
        const std::string p = dir + "/" + n;
        struct stat sb;
        if (stat(p.c_str(), &sb) != 0) continue;
        if (S_ISDIR(sb.st_mode)) {
            if (recursive) walk_dir(p, true, files);
        } else if (S_ISREG(sb.st_mode)) files.push_back(p);
    }
}

// Indexer::collect_files; throws Error(KJC_MODEL_NOT_FOUND, "Path not found: ...") for a missing input
inline std::vector<std::string> collect_files(const std::vector<std::string>& inputs, const LoaderOptions& o) {
    std::vector<std::string> files;
    for (const std::string& in : inputs) {
        struct stat sb;
        if (stat(in.c_str(), &sb) != 0) throw Error(KJC_MODEL_NOT_FOUND, "Path not found: " + in);
        if (S_ISREG(sb.st_mode)) {
            if (supported_extension(in, o)) files.push_back(in);
        } else if (S_ISDIR(sb.st_mode)) {
            std::vector<std::string> found;
            walk_dir(in, o.recursive, found);
            for (const std::string& p : found) {
                const std::string name = p.substr(p.find_last_of('/') + 1);
                if (!o.include_hidden && !name.empty() && name[0] == '.') continue;
                bool excluded = false;
                for (const std::string& pat : o.exclude_patterns)
                    if (fnmatch(pat.c_str(), p.c_str(), 0) == 0) excluded = true;
                if (excluded) continue;
                if (o.max_file_size > 0 && stat(p.c_str(), &sb) == 0 && static_cast<size_t>(sb.st_size) > o.max_file_size) continue;
                if (supported_extension(p, o)) files.push_back(p);
            }
        }
    }
    return files;
}

// ----------------------------------------------------------------- index writer
inline void write_file(const std::string& path, const std::string& data) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f || fwrite(data.data(), 1, data.size(), f) != data.size()) {
        if (f) fclose(f);
        throw Error(KJC_INFERENCE_FAILED, "cannot write " + path);
    }
    fclose(f);
}
inline void mkdirs(const std::string& path) {
    for (size_t i = 1; i <= path.size(); ++i)
        if (i == path.size() || path[i] == '/') {
            const std::string p = path.substr(0, i);
            if (mkdir(p.c_str(), 0777) != 0 && errno != EEXIST) throw Error(KJC_INFERENCE_FAILED, "cannot create directory " + p);
        }
}
inline void remove_tree(const std::string& path) {
    struct stat sb;
    if (lstat(path.c_str(), &sb) != 0) return;
    if (S_ISDIR(sb.st_mode)) {
        if (DIR* d = opendir(path.c_str())) {
            while (dirent* e = readdir(d)) {
                const std::string n = e->d_name;
                if (n != "." && n != "..") remove_tree(path + "/" + n);
            }
            closedir(d);
        }
        rmdir(path.c_str());
    } else unlink(path.c_str());
}
inline uint64_t tree_size(const std::string& path) {  // calculate_index_size: sum of the regular files below `path`
    struct stat sb;
    if (lstat(path.c_str(), &sb) != 0) return 0;
    if (S_ISREG(sb.st_mode)) return static_cast<uint64_t>(sb.st_size);
    uint64_t total = 0;
    if (S_ISDIR(sb.st_mode))
        if (DIR* d = opendir(path.c_str())) {
            while (dirent* e = readdir(d)) {
                const std::string n = e->d_name;
                if (n != "." && n != "..") total += tree_size(path + "/" + n);
            }
            closedir(d);
        }
    return total;
}
inline std::string json_quote(const std::string& s) {
    std::string o = "\"";
    for (unsigned char c : s) {
        if (c == '"') o += "\\\"";
        else if (c == '\\') o += "\\\\";
        else if (c == '\n') o += "\\n";
        else if (c == '\r') o += "\\r";
        else if (c == '\t') o += "\\t";
        else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o += static_cast<char>(c);
    }
    return o + "\"";
}

// One segment under construction (SegmentBuilder): vectors / docs / metadata streamed to temp files, BM25 in memory.
class SegmentWriter {
  public:
    SegmentWriter(const std::string& temp_dir, int dimension, size_t max_docs) : dir_(temp_dir), dim_(dimension), max_docs_(max_docs) {
        mkdirs(temp_dir);
        vfd_ = open((temp_dir + "/vectors.bin.tmp").c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0666);
        docs_ = fopen((temp_dir + "/docs.bin.tmp").c_str(), "wb");
        meta_ = fopen((temp_dir + "/metadata.jsonl.tmp").c_str(), "wb");
        if (vfd_ < 0 || !docs_ || !meta_) throw Error(KJC_INFERENCE_FAILED, "cannot create segment files in " + temp_dir);
    }
    ~SegmentWriter() {
        if (vfd_ >= 0) close(vfd_);
        if (docs_) fclose(docs_);
        if (meta_) fclose(meta_);
    }
    size_t len() const { return count_; }
    size_t room() const { return max_docs_ > count_ ? max_docs_ - count_ : 0; }
    bool full() const { return count_ >= max_docs_; }
    // text + metadata of the next document; its embedding row goes to vectors_fd() at byte offset `returned id * dim * 4`
    size_t add_text(const std::string& text, const std::vector<std::pair<std::string, std::string>>& meta) {
        const size_t id = count_++;
        offsets_.push_back(cur_);
        fwrite(text.data(), 1, text.size(), docs_);
        fputc('\n', docs_);
        cur_ += text.size() + 1;
        std::string line = "{";
        for (size_t i = 0; i < meta.size(); ++i) line += (i ? "," : "") + json_quote(meta[i].first) + ":" + json_quote(meta[i].second);
        line += "}\n";
        fwrite(line.data(), 1, line.size(), meta_);
        bm25_.add_document(id, text);
        return id;
    }
    int vectors_fd() const { return vfd_; }
    void write_rows(size_t first_id, const float* rows, size_t n) {  // straight from the caller's (pinned) buffer
        const size_t bytes = n * dim_ * sizeof(float);
        size_t done = 0;
        while (done < bytes) {
            const ssize_t w = pwrite(vfd_, reinterpret_cast<const char*>(rows) + done, bytes - done, static_cast<off_t>(first_id * dim_ * sizeof(float) + done));
            if (w <= 0) throw Error(KJC_INFERENCE_FAILED, "short write to vectors.bin");
            done += static_cast<size_t>(w);
        }
    }
    // SegmentBuilder::flush: move the temp files into place and write docs.idx, bm25.bin, segment.json
    void flush(const std::string& segment_dir, uint64_t segment_id) {
        close(vfd_); vfd_ = -1;
        fclose(docs_); docs_ = nullptr;
        fclose(meta_); meta_ = nullptr;
        mkdirs(segment_dir);
        auto mv = [&](const char* a, const char* b) {
            if (rename((dir_ + "/" + a).c_str(), (segment_dir + "/" + b).c_str()) != 0) throw Error(KJC_INFERENCE_FAILED, std::string("cannot move ") + a);
        };
        mv("vectors.bin.tmp", "vectors.bin");
        mv("docs.bin.tmp", "docs.bin");
        mv("metadata.jsonl.tmp", "metadata.jsonl");
        std::string idx;
        const uint64_t n = offsets_.size();
        idx.append(reinterpret_cast<const char*>(&n), 8);
        for (uint64_t o : offsets_) idx.append(reinterpret_cast<const char*>(&o), 8);
        write_file(segment_dir + "/docs.idx", idx);
        write_file(segment_dir + "/bm25.bin", bm25_.to_bincode());
        const uint64_t total_bytes = static_cast<uint64_t>(count_) * dim_ * 4 + cur_;
        write_file(segment_dir + "/segment.json", "{\n  \"id\": " + std::to_string(segment_id) + ",\n  \"doc_count\": " + std::to_string(count_) +
                                                      ",\n  \"dimension\": " + std::to_string(dim_) + ",\n  \"created_at\": " + std::to_string(static_cast<uint64_t>(time(nullptr))) +
                                                      ",\n  \"total_bytes\": " + std::to_string(total_bytes) + "\n}");
        rmdir(dir_.c_str());
    }

  private:
    std::string dir_;
    int dim_;
    size_t max_docs_, count_ = 0;
    int vfd_ = -1;
    FILE *docs_ = nullptr, *meta_ = nullptr;
    std::vector<uint64_t> offsets_;
    uint64_t cur_ = 0;
    Bm25Index bm25_;
};

class IndexWriter {
  public:
    // IndexWriter::open (create == true: writes config.json) / open_existing
    IndexWriter(const std::string& root, bool create, int dimension, size_t max_docs_per_segment, const std::string& embedding_model) : root_(root) {
        if (create) {
            mkdirs(root);
            mkdirs(root + "/segments");
            dim_ = dimension;
            max_docs_ = max_docs_per_segment;
            write_file(root + "/config.json", "{\n  \"dimension\": " + std::to_string(dim_) + ",\n  \"max_docs_per_segment\": " + std::to_string(max_docs_) +
                                                  ",\n  \"max_segment_memory\": 104857600,\n  \"embedding_model\": " +
                                                  (embedding_model.empty() ? std::string("null") : json_quote(embedding_model)) +
                                                  ",\n  \"model_name\": null,\n  \"created_at\": null,\n  \"version\": 1\n}");
        } else {
            const std::string txt = read_text_file(root + "/config.json", KJC_MODEL_NOT_FOUND);
            const Json cfg = JsonParser(txt.data(), txt.size()).parse();
            dim_ = static_cast<int>(cfg.number("dimension", 0));
            max_docs_ = static_cast<size_t>(cfg.number("max_docs_per_segment", 10000));
        }
        // find_next_segment_id + the document count of the segments already there
        if (DIR* d = opendir((root + "/segments").c_str())) {
            while (dirent* e = readdir(d)) {
                const std::string n = e->d_name;
                if (n.rfind("seg_", 0) != 0) continue;
                char* end = nullptr;
                const unsigned long long id = strtoull(n.c_str() + 4, &end, 10);
                if (end && *end == 0) next_id_ = std::max<uint64_t>(next_id_, id + 1);
                if (!create) {
                    const std::string mp = root + "/segments/" + n + "/segment.json";
                    struct stat sb;
                    if (stat(mp.c_str(), &sb) == 0) {
                        const std::string txt = read_text_file(mp, KJC_LOAD_FAILED);
                        total_docs_ += static_cast<size_t>(JsonParser(txt.data(), txt.size()).parse().number("doc_count", 0));
                    }
                }
            }
            closedir(d);
        }
    }
    int dimension() const { return dim_; }
    size_t len() const { return total_docs_; }
    // the segment that takes the next document (a full one is flushed first); never null
    SegmentWriter& current() {
        if (cur_ && cur_->full()) flush_current();
        if (!cur_) cur_.reset(new SegmentWriter(root_ + "/temp/seg_" + std::to_string(next_id_), dim_, max_docs_));
        return *cur_;
    }
    void note_added(size_t n) { total_docs_ += n; }
    void commit() {  // IndexWriter::commit
        flush_current();
        write_file(root_ + "/index.json", "{\n  \"total_docs\": " + std::to_string(total_docs_) + ",\n  \"segment_count\": " + std::to_string(next_id_) +
                                              ",\n  \"dimension\": " + std::to_string(dim_) + "\n}");
        remove_tree(root_ + "/temp");
    }

  private:
    void flush_current() {
        if (!cur_) return;
        if (cur_->len() > 0) {
            char name[32];
            snprintf(name, sizeof name, "seg_%06llu", static_cast<unsigned long long>(next_id_));
            cur_->flush(root_ + "/segments/" + name, next_id_);
            ++next_id_;
        }
        cur_.reset();
    }
    std::string root_;
    int dim_ = 0;
    size_t max_docs_ = 10000, total_docs_ = 0;
    uint64_t next_id_ = 0;
    std::unique_ptr<SegmentWriter> cur_;
};

}  // namespace kj
