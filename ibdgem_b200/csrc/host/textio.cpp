#include "textio.h"

#include <cstring>

namespace ibdhost {

bool ends_with_gz(const std::string &fn) {
    return fn.size() >= 3 && fn.compare(fn.size() - 3, 3, ".gz") == 0;
}

bool LineReader::open(const std::string &path) {
    close();
    path_ = path;
    if (!ends_with_gz(path)) {
        // same diagnostics as fileOpen() for a plain file that cannot be opened
        FILE *f = fopen(path.c_str(), "r");
        if (!f) {
            fprintf(stderr, "Failed to open %s.\n", path.c_str());
            perror("Error");
            return false;
        }
        fclose(f);
    }
    gz_ = gzopen(path.c_str(), "r");
    if (!gz_) return false;
    gzbuffer(gz_, 1 << 20);
    buf_.resize(1 << 22);
    beg_ = end_ = 0;
    eof_ = false;
    return true;
}

void LineReader::close() {
    if (gz_) gzclose(gz_);
    gz_ = nullptr;
}

bool LineReader::fill() {
    if (eof_) return false;
    if (beg_ > 0) {
        memmove(buf_.data(), buf_.data() + beg_, end_ - beg_);
        end_ -= beg_;
        beg_ = 0;
    }
    if (end_ == buf_.size()) buf_.resize(buf_.size() * 2);
    const int n = gzread(gz_, buf_.data() + end_, (unsigned)std::min<size_t>(buf_.size() - end_, 1u << 30));
    if (n <= 0) {
        eof_ = true;
        return false;
    }
    end_ += (size_t)n;
    return true;
}

bool LineReader::next(const char **line, size_t *len) {
    size_t scan = beg_;
    for (;;) {
        const char *nl = (const char *)memchr(buf_.data() + scan, '\n', end_ - scan);
        if (nl) {
            *line = buf_.data() + beg_;
            *len = (size_t)(nl - (buf_.data() + beg_)) + 1;
            beg_ += *len;
            return true;
        }
        scan = end_ - beg_;  // offset relative to beg_, which fill() moves to 0
        if (!fill()) {
            if (end_ > beg_) {  // last line without a newline
                *line = buf_.data() + beg_;
                *len = end_ - beg_;
                beg_ = end_;
                return true;
            }
            return false;
        }
        scan += beg_;
    }
}

}  // namespace ibdhost
