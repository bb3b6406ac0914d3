#pragma once
#include <string>
#include <stdexcept>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
namespace boost { namespace iostreams {
class mapped_file_source {
public:
    mapped_file_source() : m_data(0), m_size(0) {}
    explicit mapped_file_source(std::string const& path) : m_data(0), m_size(0) { open(path); }
    explicit mapped_file_source(const char* path) : m_data(0), m_size(0) { open(std::string(path)); }
    mapped_file_source(mapped_file_source const& o) : m_data(o.m_data), m_size(o.m_size) {}  // views share, never unmapped
    void open(std::string const& path) {
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot open " + path);
        struct stat st; fstat(fd, &st);
        m_size = size_t(st.st_size);
        if (m_size) {
            void* p = mmap(0, m_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) { ::close(fd); throw std::runtime_error("mmap failed " + path); }
            m_data = static_cast<const char*>(p);
        }
        ::close(fd);
    }
    bool is_open() const { return m_data != 0; }
    const char* data() const { return m_data; }
    size_t size() const { return m_size; }
    void close() {}
private:
    const char* m_data; size_t m_size;
};
}}
