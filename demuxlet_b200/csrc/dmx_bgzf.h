// dmx_bgzf.h -- host-side plumbing shared by the htslib-free readers (dmx_bam.cpp, dmx_vcf.cpp): a small pool of
// persistent worker threads and a byte stream over a gzip / BGZF / plain file.  A BGZF file (gzip members of <= 64 KB
// with the 'BC' extra field that gives each block's compressed size, SAM spec 4.1 -- what bgzip and samtools write) is
// mapped and inflated block-parallel, one batch of blocks ahead of the consumer -- htslib's bgzf_mt idea; any other
// gzip file (or a plain one) goes through zlib's gzread.  Checked on generated BGZF files against the gzread path and
// python's gzip module (tests/test_bam.py, tests/test_vcf.py); the reference holds no BGZF file (its tutorial VCF is
// plain gzip), so the block framing is "unpinned" against htslib's own writer.
#pragma once
#include <fcntl.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "dmx_inflate.h"

namespace dmx_io {

// host threads to use: the CPUs this process may run on (<= 32), or the value of the environment variable `env_name`
inline int host_threads(const char* env_name) {
  int n = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::max(1, std::min(n, CPU_COUNT(&set)));
  if (const char* env = env_name ? getenv(env_name) : nullptr) n = std::max(1, std::min(atoi(env), 256));
  return n;
}

// Host threads that live as long as one sweep: run(nt, f) calls f(t) for t in [0, nt), t = 0 on the calling thread.
class Pool {
 public:
  explicit Pool(int n) : n_(std::max(1, n)) {
    for (int t = 1; t < n_; ++t) th_.emplace_back([this, t]() { loop(t); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& x : th_) x.join();
  }
  int size() const { return n_; }
  void run(int nt, const std::function<void(int)>& f) {
    nt = std::min(nt, n_);
    if (nt <= 1) { f(0); return; }
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &f;
      nt_ = nt;
      pending_ = nt - 1;
      ++gen_;
    }
    cv_.notify_all();
    f(0);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this]() { return pending_ == 0; });
  }

 private:
  void loop(int t) {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(m_);
    for (;;) {
      cv_.wait(lk, [&]() { return gen_ != seen; });
      seen = gen_;
      if (stop_) return;
      if (t >= nt_) continue;
      const std::function<void(int)>* f = fn_;
      lk.unlock();
      (*f)(t);
      lk.lock();
      if (--pending_ == 0) done_.notify_one();
    }
  }
  int n_, nt_ = 0, pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
  const std::function<void(int)>* fn_ = nullptr;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> th_;
};

// Decoded BAM bytes as a sliding window.  more(keep_from) drops the bytes before keep_from, appends the next batch and
// returns false when there is nothing left (the window then holds only the kept tail).
class ByteStream {
 public:
  ~ByteStream() {
    if (next_.valid()) next_.wait();
    if (gz_) gzclose(gz_);
    if (map_ && map_ != MAP_FAILED) munmap((void*)map_, map_size_);
  }
  bool open(const char* path, int nth) {
    pool_.reset(new Pool(nth));
    // test knobs: small windows exercise the carry-over of incomplete records / lines
    if (const char* env = getenv("DMX_IO_BATCH_BLOCKS")) batch_blocks_ = (size_t)std::max(1, atoi(env));
    if (const char* env = getenv("DMX_IO_CHUNK_BYTES")) gz_chunk_ = (size_t)std::max(1, atoi(env));
    if (!strcmp(path, "-")) {          // standard input, as htslib reads it
      gz_ = gzdopen(dup(0), "rb");
      if (gz_) gzbuffer(gz_, 1 << 20);
      return gz_ != nullptr;
    }
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    uint8_t h[18];
    const ssize_t got = pread(fd, h, 18, 0);
    const bool bgzf = got == 18 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) && h[10] == 6 && h[11] == 0 &&
                      h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0;
    struct stat st;
    if (!bgzf || fstat(fd, &st) != 0 || st.st_size <= 0) {
      ::close(fd);
      gz_ = gzopen(path, "rb");
      if (gz_) gzbuffer(gz_, 1 << 20);
      return gz_ != nullptr;
    }
    map_size_ = (size_t)st.st_size;
    map_ = (const uint8_t*)mmap(nullptr, map_size_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map_ == MAP_FAILED) { map_ = nullptr; return false; }
    madvise((void*)map_, map_size_, MADV_SEQUENTIAL);
    parallel_ = true;
    schedule();
    return true;
  }
  const uint8_t* data() const { return win_.data() + beg_; }
  size_t size() const { return win_.size() - beg_; }
  bool bad() const { return bad_; }
  // share of the compressed file whose bytes have been handed out or are being inflated (0: unknown)
  double decoded_share() const { return map_size_ ? (double)sched_off_ / (double)map_size_ : 0.0; }
  bool more(size_t keep_from) {
    const size_t keep = size() - keep_from;
    static const uint8_t kNone = 0;
    const uint8_t* tail = keep ? data() + keep_from : &kNone;
    if (!parallel_) {
      if (gz_eof_) { if (keep) memmove(win_.data(), tail, keep); win_.resize(keep); beg_ = 0; return false; }
      std::vector<uint8_t> nw(keep + gz_chunk_);
      if (keep) memcpy(nw.data(), tail, keep);
      size_t got = 0;
      while (got < gz_chunk_) {
        const int k = gzread(gz_, nw.data() + keep + got, (unsigned)(gz_chunk_ - got));
        if (k < 0) { bad_ = true; break; }
        if (k == 0) { gz_eof_ = true; break; }
        got += (size_t)k;
      }
      nw.resize(keep + got);
      win_.swap(nw);
      beg_ = 0;
      return got > 0 && !bad_;
    }
    for (;;) {
      if (!next_.valid()) { if (keep) memmove(win_.data(), tail, keep); win_.resize(keep); beg_ = 0; return false; }
      if (!next_.get()) { bad_ = true; return false; }
      // ahead_ was inflated behind kSlack spare bytes: the kept tail goes in front of it without moving the batch
      if (keep <= kSlack) {
        if (keep) memcpy(ahead_.data() + kSlack - keep, tail, keep);
        win_.swap(ahead_);
        beg_ = kSlack - keep;
      } else {
        std::vector<uint8_t> nw(keep + ahead_.size() - kSlack);
        memcpy(nw.data(), tail, keep);
        memcpy(nw.data() + keep, ahead_.data() + kSlack, ahead_.size() - kSlack);
        win_.swap(nw);
        beg_ = 0;
      }
      const bool grew = size() > keep;
      schedule();
      if (grew) return true;
      // a batch of empty blocks (the end-of-file marker): keep going
      return more(0);
    }
  }

 private:
  struct Blk { size_t off, size; uint32_t isize; };
  static constexpr size_t kBatch = 256;         // blocks per batch (<= 16 MB of BAM)
  static constexpr size_t kSlack = 1 << 16;     // room in front of a batch for the unconsumed tail of the last one
  static constexpr size_t kGzChunk = 32u << 20;
  // every block starts with the 18-byte header checked in open(); BSIZE = total block size - 1
  bool index_batch(std::vector<Blk>* blocks) {
    while (blocks->size() < batch_blocks_ && file_off_ < map_size_) {
      if (file_off_ + 28 > map_size_) return false;
      const uint8_t* b = map_ + file_off_;
      if (!(b[0] == 0x1f && b[1] == 0x8b && b[12] == 'B' && b[13] == 'C')) return false;
      const size_t bsize = (size_t)(b[16] | (b[17] << 8)) + 1;
      if (bsize < 26 || file_off_ + bsize > map_size_) return false;
      uint32_t isize;
      memcpy(&isize, b + bsize - 4, 4);
      blocks->push_back(Blk{file_off_, bsize, isize});
      file_off_ += bsize;
    }
    return true;
  }
  bool inflate_batch(const std::vector<Blk>& blocks, std::vector<uint8_t>* out) {
    const size_t nb = blocks.size();
    std::vector<size_t> at(nb + 1, kSlack);
    for (size_t i = 0; i < nb; ++i) at[i + 1] = at[i] + blocks[i].isize;
    out->resize(at.back());
    std::atomic<size_t> nextb{0};
    std::atomic<bool> ok{true};
    const bool zlib_only = getenv("DMX_IO_ZLIB") != nullptr;      // A / B switch: every block through zlib
    auto work = [&](int) {
      // every block first through the block decoder of dmx_inflate.h; whatever it cannot vouch for -- wrong length, no
      // clean end of stream, CRC-32 mismatch -- goes through zlib, whose verdict (with the same CRC check) stands
      std::unique_ptr<FastInflate> fast(zlib_only ? nullptr : new FastInflate());
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      if (inflateInit2(&zs, -15) != Z_OK) { ok = false; return; }
      for (;;) {
        const size_t i = nextb.fetch_add(1);
        if (i >= nb) break;
        const Blk& bl = blocks[i];
        if (bl.isize == 0) continue;
        uint32_t crc;
        memcpy(&crc, map_ + bl.off + bl.size - 8, 4);
        uint8_t* dst = out->data() + at[i];
        if (fast && fast->run(map_ + bl.off + 18, bl.size - 26, dst, bl.isize) &&
            (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, bl.isize) == crc)
          continue;
        inflateReset(&zs);
        zs.next_in = const_cast<Bytef*>(map_ + bl.off + 18);
        zs.avail_in = (uInt)(bl.size - 26);
        zs.next_out = dst;
        zs.avail_out = bl.isize;
        const int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.avail_out != 0 || (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, bl.isize) != crc) ok = false;
      }
      inflateEnd(&zs);
    };
    pool_->run((int)std::min<size_t>((size_t)pool_->size(), std::max<size_t>(nb, 1)), work);
    return ok;
  }
  void schedule() {      // start inflating the next batch in the background
    sched_off_ = file_off_;
    if (file_off_ >= map_size_ || index_bad_) return;
    next_ = std::async(std::launch::async, [this]() {
      std::vector<Blk> blocks;
      blocks.reserve(kBatch);
      if (!index_batch(&blocks)) { index_bad_ = true; return false; }
      return inflate_batch(blocks, &ahead_);
    });
  }
  gzFile gz_ = nullptr;
  bool parallel_ = false, bad_ = false, gz_eof_ = false, index_bad_ = false;
  std::unique_ptr<Pool> pool_;
  const uint8_t* map_ = nullptr;
  size_t map_size_ = 0, file_off_ = 0, sched_off_ = 0, batch_blocks_ = kBatch, gz_chunk_ = kGzChunk;
  std::vector<uint8_t> win_, ahead_;
  size_t beg_ = 0;
  std::future<bool> next_;
};

}  // namespace dmx_io
