// tests/hosttool/hosttool.cc -- TEST TOOL: drives the host mirror's file classes (kbbq_b200/host) without a GPU.
//   hosttool load FILE                     parse only (timing)
//   hosttool dump FILE [use_oq]            one line per record: name \t seq \t qual+33 \t rg-id \t second
//   hosttool rewrite FILE QUALS OUT [set_oq] [use_oq]
//        QUALS holds one line of new qualities (+33, read orientation) per record; every record is passed through
//        HTSFile::next / recalibrate / write exactly as recalibrate_and_write's generic loop does
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>

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
