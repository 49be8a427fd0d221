// FASTA/FASTQ record reader with the record semantics of the reference's parser (Heng Li's
// kseq.h as used by Shark: kseq.h:177-218), written from those semantics over a large zlib
// buffer.  What must match:
//   * a record starts at the next '>' or '@' byte (anywhere, when the previous record ended at a
//     quality string; at a line start otherwise), name = bytes up to the first isspace(), the
//     rest of the header line is dropped;
//   * sequence = the following lines joined, until a line that starts with '>', '@' or '+';
//     empty lines are skipped; a trailing '\r' is stripped from every line that leaves more
//     than one byte accumulated;
//   * after a '+' line, quality lines are appended until the quality is at least as long as
//     the sequence; a different length is an error (-2) after which the NEXT call resynchronises
//     at the next '>'/'@' byte - the reader stays usable, which the reference's batching relies
//     on (FastqSplitter.hpp:53,61: a failed read only ends the current batch);
//   * input may be gzip or plain (gzread).
#pragma once
#include <zlib.h>

#include <cctype>
#include <cstring>
#include <string>

namespace shkhost {

class FastxReader {
public:
    explicit FastxReader(const char *path, size_t buf_size = 1u << 22) : buf_(new unsigned char[buf_size]), cap_(buf_size)
    {
        f_ = gzopen(path, "r");
        if (f_) gzbuffer(f_, 1u << 20);
    }
    ~FastxReader()
    {
        if (f_) gzclose(f_);
        delete[] buf_;
    }
    FastxReader(const FastxReader &) = delete;
    FastxReader &operator=(const FastxReader &) = delete;
    bool ok() const { return f_ != nullptr; }

    // Reads the next record into name/seq/qual (qual empty for FASTA records).
    // Returns the sequence length (>= 0), -1 at end of file, -2 for a truncated quality string,
    // -3 for a stream error - the values of kseq_read.
    long read(std::string &name, std::string &seq, std::string &qual)
    {
        int c;
        if (last_char_ == 0) {  // jump to the next header byte
            while ((c = getc()) >= 0 && c != '>' && c != '@') {
            }
            if (c < 0) return c;
            last_char_ = c;
        }
        seq.clear();
        qual.clear();
        int delim = 0;
        long r = get_until(kSpace, name, false, &delim);
        if (r < 0) return r;
        if (delim != '\n') {
            scratch_.clear();
            get_until(kLine, scratch_, false, nullptr);  // comment: dropped
        }
        while ((c = getc()) >= 0 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;
            seq.push_back((char)c);
            get_until(kLine, seq, true, nullptr);
        }
        if (c == '>' || c == '@') last_char_ = c;
        if (c != '+') return (long)seq.size();  // FASTA record
        while ((c = getc()) >= 0 && c != '\n') {
        }
        if (c == -1) return -2;
        while (get_until(kLine, qual, true, nullptr) >= 0 && qual.size() < seq.size()) {
        }
        last_char_ = 0;
        if (seq.size() != qual.size()) return -2;
        return (long)seq.size();
    }

private:
    enum Delim { kSpace, kLine };

    bool fill()
    {
        if (eof_ || err_) return false;
        int n = f_ ? gzread(f_, buf_, (unsigned)cap_) : 0;
        begin_ = 0;
        if (n <= 0) {
            end_ = 0;
            if (n < 0) err_ = true;
            eof_ = true;
            return false;
        }
        end_ = (size_t)n;
        return true;
    }
    int getc()
    {
        if (err_) return -3;
        if (begin_ >= end_ && !fill()) return err_ ? -3 : -1;
        return (int)buf_[begin_++];
    }
    // Appends bytes up to (not including) the delimiter; consumes the delimiter.  Returns the
    // accumulated length, or -1 when nothing at all could be read (end of file).
    long get_until(Delim d, std::string &out, bool append, int *dret)
    {
        if (!append) out.clear();
        if (dret) *dret = 0;
        bool got_any = false;
        for (;;) {
            if (err_) return -3;
            if (begin_ >= end_ && !fill()) break;
            got_any = true;
            const unsigned char *p = buf_ + begin_;
            const size_t avail = end_ - begin_;
            size_t i = 0;
            if (d == kLine) {
                const void *q = memchr(p, '\n', avail);
                i = q ? (size_t)((const unsigned char *)q - p) : avail;
            } else {
                while (i < avail && !isspace(p[i])) ++i;
            }
            out.append((const char *)p, i);
            begin_ += i;
            if (i < avail) {
                if (dret) *dret = p[i];
                ++begin_;
                break;
            }
        }
        if (!got_any) return -1;
        if (d == kLine && out.size() > 1 && out.back() == '\r') out.pop_back();
        return (long)out.size();
    }

    gzFile f_ = nullptr;
    unsigned char *buf_;
    size_t cap_, begin_ = 0, end_ = 0;
    bool eof_ = false, err_ = false;
    int last_char_ = 0;
    std::string scratch_;
};

}  // namespace shkhost
