// Signature -- host mirror of the reference's Signature type and its JSON form
// (src/lib.rs:546-675, serde derive + the hand-written KmerMinHash Serialize/Deserialize at
// src/lib.rs:62-139).  Formatting only: no sketch arithmetic happens here.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "minhash.hpp"

namespace smb200 {

struct Signature {
    // field order == JSON field order (lib.rs:546-565)
    std::string class_ = "sourmash_signature";  // default_class, lib.rs:571-573
    std::string email = "";
    std::string hash_function = "0.murmur64";   // Default impl, lib.rs:648-661
    bool has_filename = false;
    std::string filename;
    bool has_name = false;
    std::string name;
    std::string license = "CC0";                 // default_license, lib.rs:567-569
    std::vector<std::unique_ptr<KmerMinHash>> signatures;
    double version = 0.4;                        // default_version, lib.rs:575-577

    std::mutex mu;  // as KmerMinHash::mu; covers the sketches the signature owns

    Signature() = default;
    Signature *clone_meta() const;  // everything but the sketches
    bool equals(Signature &other);  // lib.rs:663-675
    void to_json(std::string &out);  // serde_json::to_string(&Signature), ffi.rs:498
};

// serde_json::to_string(&Vec<&Signature>) (ffi.rs:531) into a malloc'd buffer of *len bytes (not NUL-terminated)
char *signatures_to_json(Signature *const *sigs, size_t n, size_t *len);
// Signature::load_signatures (lib.rs:593-645): parse a JSON array of signatures, flatten to one
// Signature per sketch, keep those passing the ksize / moltype filter.  Throws SourmashError.
std::vector<std::unique_ptr<Signature>> load_signatures(const char *data, size_t len, size_t ksize,
                                                        const char *moltype /*nullable*/);
std::vector<std::unique_ptr<Signature>> load_signatures_path(const char *path, size_t ksize, const char *moltype);

}  // namespace smb200
