// tests/hosttool/hosttool.cc -- TEST TOOL: drives the host mirror's file classes (kbbq_b200/host) without a GPU.
//   hosttool load FILE                     parse only (timing)
//   hosttool dump FILE [use_oq]            one line per record: name \t seq \t qual+33 \t rg-id \t second
//   hosttool rewrite FILE QUALS OUT [set_oq] [use_oq]
//        QUALS holds one line of new qualities (+33, read orientation) per record; every record is passed through
//        HTSFile::next / recalibrate / write exactly as recalibrate_and_write's generic loop does
//   hosttool prepack FILE K NGPUS
//        a background Session over NGPUS devices (its contexts come up, or fail, on a side thread) packs every GPU's first
//        batch without touching a context: prints one line per batch (lo hi gpu n_words slots prepacked), then whether
//        the contexts arrived
//   hosttool deflate FILE OUT BLOCK MINMATCH
//        the sink's own DEFLATE encoder (fast_deflate.hh) on FILE cut into BLOCK-byte pieces; OUT = per piece a
//        little-endian u32 stream length followed by the raw stream
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <memory>
#include <vector>

#include "../../kbbq_b200/host/fast_deflate.hh"
#include "../../kbbq_b200/host/kbbq_host_api.hh"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string mode = argv[1], fn = argv[2];
    bool set_oq = false, use_oq = false;
    const int nthreads = getenv("HOSTTOOL_THREADS") ? atoi(getenv("HOSTTOOL_THREADS")) : 0;  // 0 = the library default
    for (int i = 3; i < argc; ++i) {
        if (!strcmp(argv[i], "set_oq")) set_oq = true;
        if (!strcmp(argv[i], "use_oq")) use_oq = true;
    }
    if (mode == "deflate") {
        if (argc < 6) return 2;
        std::ifstream in(fn, std::ios::binary);
        const std::string data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        std::ofstream out(argv[3], std::ios::binary);
        const size_t block = (size_t)atol(argv[4]);
        const int min_match = atoi(argv[5]);
        std::vector<uint8_t> z(kbbq::fdeflate::bound(block));
        size_t p = 0;
        do {  // an empty file is one empty piece
            const size_t n = std::min(block, data.size() - p);
            const uint32_t c = (uint32_t)kbbq::fdeflate::compress((const uint8_t*)data.data() + p, n, z.data(), z.size(), min_match);
            if (c == 0) return 6;
            out.write((const char*)&c, 4);
            out.write((const char*)z.data(), c);
            p += n;
        } while (p < data.size());
        return 0;
    }
    try {
        std::unique_ptr<htsiter::StoreFile> f;
        if (htsiter::is_bam_file(fn)) f.reset(new htsiter::BamFile(fn, nthreads, use_oq, set_oq));
        else f.reset(new htsiter::FastqFile(fn, nthreads));
        if (mode == "load") {  // parse only: prints the number of records and bases
            std::cout << f->store().n() << '\t' << f->store().off.back() << '\n';
        } else if (mode == "dump") {
            while (f->next() >= 0) {
                readutils::CReadData r = f->get();
                std::string q;
                for (uint8_t c : r.qual) q.push_back((char)(c + 33));
                std::cout << r.name << '\t' << r.seq << '\t' << q << '\t' << readutils::CReadData::rg_to_int[r.rg] << '\t' << (int)r.second << '\n';
            }
            for (uint64_t l : f->store().ref_len) std::cout << "#ref\t" << l << '\n';
        } else if (mode == "prepack") {
            if (argc < 5) return 2;
            const int k = atoi(argv[3]), ngpus = atoi(argv[4]);
            std::vector<int> devices;
            for (int d = 0; d < ngpus; ++d) devices.push_back(d);
            const int n_rg = (int)std::max<size_t>(1, readutils::CReadData::rg_to_int.size());
            kbbq::Session ses(k, n_rg, (int)std::max<uint32_t>(1, f->store().max_len), devices, true);
            ses.prepack(f->store());
            for (const kbbq::SessionBatch& b : ses.batches())
                std::cout << b.lo << '\t' << b.hi << '\t' << b.gpu << '\t' << b.n_words << '\t' << b.slots << '\t' << (b.prepacked ? 1 : 0) << '\n';
            try {
                const size_t n_ctx = ses.contexts().size();
                std::cout << "contexts\t" << n_ctx << '\n';
            } catch (const std::exception& e) {
                std::cout << "contexts\tfailed\t" << e.what() << '\n';
            }
        } else if (mode == "rewrite") {
            std::ifstream qin(argv[3]);
            if (f->open_out(argv[4]) < 0) return 3;
            std::string line;
            while (f->next() >= 0) {
                if (!std::getline(qin, line)) return 4;
                std::vector<uint8_t> q;
                for (char c : line) q.push_back((uint8_t)(c - 33));
                f->recalibrate(q);
                if (f->write() < 0) return 5;
            }
        } else {
            return 2;
        }
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
