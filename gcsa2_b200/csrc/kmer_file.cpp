/*
  kmer_file.cpp -- reads the reference's construction input on the host: kmer files (.graph binary, .gcsa2 text)
  and NodeMapping files.

  Replaces readBinary / readText (src/files.cpp:86-167), KMer(tokens, alpha, successor) and KMer::chars
  (src/support.cpp:620-635), Key::encode (include/gcsa/support.h:385-396), Node::encode(token)
  (src/support.cpp:565-592) and NodeMapping::load (src/support.cpp:335-342).  The records feed
  gcsa_b200_build_from_kmers[_mapped] and gcsa_b200_verify_index[_mapped].

  Formats
    .graph   sections of { u64 flags (0), u64 kmer_count, u64 kmer_length } followed by kmer_count records
             { u64 key, u64 from, u64 to } (include/gcsa/files.h:40-52, support.h:475-497);
    .gcsa2   one kmer per line, five tab-separated columns: the kmer, its start position "id:offset" (or
             "id:-offset" on the reverse strand), the predecessor characters and the successor characters (comma
             separated), the successor positions (comma separated; one record per successor position);
    mapping  u64 first_node, u64 next_node, (next_node - first_node) u64 node ids.

  Where the reference exits the process (bad flags, mixed kmer lengths, truncated file) this returns
  GCSA_B200_ERR_INVALID with a message; a text line without five columns is skipped, as there.
*/
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gcsa2_b200.h"
#include "internal.h"

namespace
{

typedef uint64_t u64;
constexpr u64 MAX_KMER_LENGTH = 16;          // Key::MAX_LENGTH, support.h:382
constexpr u64 OFFSET_MASK = 0x3FF;           // Node::OFFSET_MASK, support.h:446-449
constexpr u64 ORIENTATION_MASK = 0x400;
constexpr u64 ID_OFFSET = 11;

struct Failure { std::string what; };

struct Records
{
  std::vector<u64> key, from, to;
  u64 kmer_length = ~(u64)0;
};

void setLength(Records& r, u64 length, const std::string& where)
{
  if(r.kmer_length == ~(u64)0)
  {
    if(length == 0 || length > MAX_KMER_LENGTH) { throw Failure{where + ": invalid kmer length " + std::to_string(length)}; }
    r.kmer_length = length;
  }
  else if(length != r.kmer_length)
  {
    throw Failure{where + ": kmer length " + std::to_string(length) + " (expected " + std::to_string(r.kmer_length) + ")"};
  }
}

// readBinary, src/files.cpp:127-167
void readBinary(const char* path, Records& r)
{
  FILE* f = std::fopen(path, "rb");
  if(f == nullptr) { throw Failure{std::string("cannot open ") + path}; }
  struct Close { FILE* f; ~Close() { std::fclose(f); } } close = { f };
  for(u64 section = 0; ; section++)
  {
    u64 header[3];
    size_t got = std::fread(header, 1, sizeof(header), f);
    if(got == 0) { break; }
    if(got != sizeof(header)) { throw Failure{std::string(path) + ": truncated header in section " + std::to_string(section)}; }
    if(header[0] != 0) { throw Failure{std::string(path) + ": invalid flags in section " + std::to_string(section)}; }
    setLength(r, header[2], std::string(path) + ", section " + std::to_string(section));
    const u64 count = header[1];
    std::vector<u64> buffer(3 * (size_t)std::min<u64>(count, 1u << 20));
    for(u64 done = 0; done < count; )
    {
      u64 m = std::min<u64>(count - done, 1u << 20);
      if(std::fread(buffer.data(), 24, m, f) != m) { throw Failure{std::string(path) + ": unexpected end of file"}; }
      for(u64 i = 0; i < m; i++) { r.key.push_back(buffer[3 * i]); r.from.push_back(buffer[3 * i + 1]); r.to.push_back(buffer[3 * i + 2]); }
      done += m;
    }
  }
}

std::vector<std::string> split(const std::string& s, char sep)
{
  // std::getline semantics: "a,,b" -> a, "", b; a trailing separator adds nothing; "" -> nothing
  std::vector<std::string> out;
  size_t start = 0;
  while(start < s.size())
  {
    size_t end = s.find(sep, start);
    if(end == std::string::npos) { out.push_back(s.substr(start)); break; }
    out.push_back(s.substr(start, end - start));
    start = end + 1;
  }
  return out;
}

// Node::encode(token), src/support.cpp:565-592; an invalid token is node 0, as there
u64 encodeNode(const std::string& token)
{
  char* end = nullptr;
  const char* begin = token.c_str();
  u64 id = std::strtoull(begin, &end, 10);
  size_t separator = (size_t)(end - begin);
  if(end == begin || separator + 1 >= token.size()) { return 0; }
  bool reverse = false;
  if(token[separator + 1] == '-') { reverse = true; separator++; }
  if(separator + 1 >= token.size()) { return 0; }
  u64 offset = std::strtoull(begin + separator + 1, nullptr, 10);
  if(offset > OFFSET_MASK) { return 0; }
  return (id << ID_OFFSET) | offset | (reverse ? ORIENTATION_MASK : 0);
}

// KMer::chars, src/support.cpp:629-635: every other character of a comma-separated list
u64 charSet(const std::string& token, const uint8_t* char2comp)
{
  u64 value = 0;
  for(size_t i = 0; i < token.size(); i += 2) { value |= (u64)1 << char2comp[(unsigned char)token[i]]; }
  return value & 0xFF;
}

// readText, src/files.cpp:86-124
void readText(const char* path, const uint8_t* char2comp, Records& r)
{
  FILE* f = std::fopen(path, "r");
  if(f == nullptr) { throw Failure{std::string("cannot open ") + path}; }
  struct Close { FILE* f; ~Close() { std::fclose(f); } } close = { f };
  std::string line;
  char chunk[1 << 16];
  while(true)
  {
    line.clear();
    bool any = false, complete = false;
    while(std::fgets(chunk, sizeof(chunk), f) != nullptr)
    {
      any = true; line += chunk;
      if(!line.empty() && line.back() == '\n') { line.pop_back(); complete = true; break; }
    }
    if(!any) { break; }
    (void)complete;
    std::vector<std::string> tokens = split(line, '\t');
    if(tokens.size() != 5) { continue; }                                  // tokenize(): reported and skipped
    setLength(r, tokens[0].size(), std::string(path));
    u64 label = 0;
    for(char c : tokens[0]) { label = (label << 3) | char2comp[(unsigned char)c]; }
    u64 key = (((label << 8) | charSet(tokens[2], char2comp)) << 8) | charSet(tokens[3], char2comp);   // Key::encode
    u64 from = encodeNode(tokens[1]);
    for(const std::string& successor : split(tokens[4], ','))
    {
      r.key.push_back(key); r.from.push_back(from); r.to.push_back(encodeNode(successor));
    }
  }
}

u64* release(std::vector<u64>& v)
{
  u64* p = (u64*)std::malloc(sizeof(u64) * (v.size() + 1));
  if(p != nullptr && !v.empty()) { std::memcpy(p, v.data(), sizeof(u64) * v.size()); }
  std::vector<u64>().swap(v);
  return p;
}

template<class Work> int guarded(const char* what, Work work)
{
  try { work(); }
  catch(const Failure& f) { gcsa_b200_internal_set_error((std::string(what) + ": " + f.what).c_str()); return GCSA_B200_ERR_INVALID; }
  catch(const std::bad_alloc&) { gcsa_b200_internal_set_error((std::string(what) + ": out of memory").c_str()); return GCSA_B200_ERR_NOMEM; }
  return 0;
}

} // namespace

extern "C" {

int gcsa_b200_read_kmer_files(const char* const* paths, int count, int binary, const uint8_t* char2comp,
                              gcsa_b200_kmers* result, int* kmer_length)
{
  if(paths == nullptr || count < 1 || result == nullptr) { return GCSA_B200_ERR_INVALID; }
  std::memset(result, 0, sizeof(*result));
  uint8_t default_table[256];
  if(char2comp == nullptr) { gcsa_b200_default_char2comp(default_table); char2comp = default_table; }
  Records r;
  int rc = guarded("gcsa_b200_read_kmer_files", [&]()
  {
    for(int i = 0; i < count; i++)
    {
      if(paths[i] == nullptr) { throw Failure{"null path"}; }
      if(binary) { readBinary(paths[i], r); } else { readText(paths[i], char2comp, r); }
    }
  });
  if(rc != 0) { return rc; }
  result->n = r.key.size();
  result->key = release(r.key); result->from = release(r.from); result->to = release(r.to);
  if(result->key == nullptr || result->from == nullptr || result->to == nullptr)
  {
    gcsa_b200_kmers_free(result);
    gcsa_b200_internal_set_error("gcsa_b200_read_kmer_files: out of memory");
    return GCSA_B200_ERR_NOMEM;
  }
  if(kmer_length != nullptr) { *kmer_length = (r.kmer_length == ~(u64)0 ? 0 : (int)r.kmer_length); }
  return 0;
}

int gcsa_b200_load_node_mapping(const char* path, uint64_t* first_node, uint64_t** ids, uint64_t* size)
{
  if(path == nullptr || first_node == nullptr || ids == nullptr || size == nullptr) { return GCSA_B200_ERR_INVALID; }
  *first_node = 0; *ids = nullptr; *size = 0;
  return guarded("gcsa_b200_load_node_mapping", [&]()
  {
    FILE* f = std::fopen(path, "rb");
    if(f == nullptr) { throw Failure{std::string("cannot open ") + path}; }
    struct Close { FILE* f; ~Close() { std::fclose(f); } } close = { f };
    u64 header[2];
    if(std::fread(header, 1, sizeof(header), f) != sizeof(header)) { throw Failure{std::string(path) + ": truncated header"}; }
    if(header[1] < header[0]) { throw Failure{std::string(path) + ": next_node < first_node"}; }
    u64 n = header[1] - header[0];
    u64* p = (u64*)std::malloc(sizeof(u64) * (n + 1));
    if(p == nullptr) { throw std::bad_alloc(); }
    if(n > 0 && std::fread(p, sizeof(u64), n, f) != n) { std::free(p); throw Failure{std::string(path) + ": unexpected end of file"}; }
    *first_node = header[0]; *ids = p; *size = n;
  });
}

} // extern "C"
