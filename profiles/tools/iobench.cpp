// Host I/O micro-benchmark (measurement only): how fast can FASTQ-sized output be written on this box?
//   iobench DIR GB THREADS  ->  write(2) sequential, pwrite from T threads, memcpy into a shared file mapping from T threads
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv)
{
    const std::string dir = argc > 1 ? argv[1] : "/dev/shm";
    const size_t total = (size_t)(atof(argc > 2 ? argv[2] : "2") * (1u << 30));
    const int T = argc > 3 ? atoi(argv[3]) : 8;
    const size_t piece = 8u << 20;
    std::vector<char> src(piece);
    for (size_t i = 0; i < piece; ++i) src[i] = (char)(i * 131 >> 3);
    const std::string path = dir + "/iobench.tmp";
    {
        int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
        double t = now();
        for (size_t o = 0; o < total; o += piece) if (write(fd, src.data(), piece) != (ssize_t)piece) return 1;
        printf("%s write(2) 1 thread: %.2f GB/s\n", dir.c_str(), total / (now() - t) / 1e9);
        close(fd);
    }
    for (int threads : {1, T / 2, T}) {
        if (threads < 1) continue;
        int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
        double t = now();
        std::vector<std::thread> th;
        for (int i = 0; i < threads; ++i)
            th.emplace_back([&, i] {
                for (size_t o = (size_t)i * piece; o < total; o += (size_t)threads * piece)
                    if (pwrite(fd, src.data(), piece, (off_t)o) != (ssize_t)piece) return;
            });
        for (auto &x : th) x.join();
        printf("%s pwrite %d threads: %.2f GB/s\n", dir.c_str(), threads, total / (now() - t) / 1e9);
        close(fd);
    }
    for (int threads : {1, T / 2, T}) {
        if (threads < 1) continue;
        int fd = open(path.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0666);
        double t = now();
        if (ftruncate(fd, (off_t)total) != 0) return 1;
        char *m = (char *)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) return 1;
        std::vector<std::thread> th;
        for (int i = 0; i < threads; ++i)
            th.emplace_back([&, i] {
                for (size_t o = (size_t)i * piece; o < total; o += (size_t)threads * piece) memcpy(m + o, src.data(), piece);
            });
        for (auto &x : th) x.join();
        munmap(m, total);
        printf("%s mmap+memcpy %d threads: %.2f GB/s\n", dir.c_str(), threads, total / (now() - t) / 1e9);
        close(fd);
    }
    // several files at once (ssv / out1 / out2 are different files): 3 writers
    {
        double t = now();
        std::vector<std::thread> th;
        for (int f = 0; f < 3; ++f)
            th.emplace_back([&, f] {
                const std::string p = path + std::to_string(f);
                int fd = open(p.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
                for (size_t o = 0; o < total / 3; o += piece) if (write(fd, src.data(), piece) != (ssize_t)piece) return;
                close(fd);
                unlink(p.c_str());
            });
        for (auto &x : th) x.join();
        printf("%s write(2) to 3 files at once: %.2f GB/s in total\n", dir.c_str(), total / (now() - t) / 1e9);
    }
    // reading: memchr over a mapped file from T threads
    {
        int fd = open(path.c_str(), O_RDONLY);
        char *m = (char *)mmap(nullptr, total, PROT_READ, MAP_PRIVATE, fd, 0);
        for (int threads : {1, T}) {
            double t = now();
            std::vector<std::thread> th;
            std::vector<size_t> cnt((size_t)threads, 0);
            for (int i = 0; i < threads; ++i)
                th.emplace_back([&, i] {
                    for (size_t o = (size_t)i * piece; o < total; o += (size_t)threads * piece) {
                        const char *p = m + o, *e = p + piece;
                        while ((p = (const char *)memchr(p, '\n', (size_t)(e - p))) != nullptr) ++cnt[(size_t)i], ++p;
                    }
                });
            for (auto &x : th) x.join();
            printf("%s mmap read + memchr %d threads: %.2f GB/s\n", dir.c_str(), threads, total / (now() - t) / 1e9);
        }
        munmap(m, total);
        close(fd);
    }
    unlink(path.c_str());
    return 0;
}
