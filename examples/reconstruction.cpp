// examples/reconstruction.cpp — the flow of the reference's examples/reconstruction.rs (and, with --partition-only,
// examples/partition.rs) written against include/soundsym.hpp:
//   -s source.wav  -t target.wav  -o out.wav  -m model.bin  [--depth 3] [--threshold 4] [--dtw] [--partition-only]
// Partitioner::from_path(source).threshold(4).depth(3) -> train (model supplied) -> partition -> SoundDictionary::from_segments
// -> partition the target with the source's model -> clone_from_dictionary -> to_sound -> write_file.
// Prints one JSON line with the segment counts, the match indices and checksums so that a test can compare it with the
// golden fixtures.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "soundsym.hpp"

using namespace soundsym;

int main(int argc, char** argv) {
    std::string src, tgt, out, model_path;
    size_t depth = 3, threshold = 4;  // examples/reconstruction.rs:44-45
    bool dtw = false, partition_only = false;
    std::string dict_dir, query_dir, splits_dir;  // --matcher <dictionary dir> <query dir>: the flow of examples/matcher.rs
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? std::string(argv[++i]) : std::string(); };
        if (a == "-s") src = next();
        else if (a == "-t") tgt = next();
        else if (a == "-o") out = next();
        else if (a == "-m") model_path = next();
        else if (a == "--depth") depth = std::stoul(next());
        else if (a == "--threshold") threshold = std::stoul(next());
        else if (a == "--dtw") dtw = true;
        else if (a == "--partition-only") partition_only = true;
        else if (a == "--matcher") dict_dir = next(), query_dir = next();
        else if (a == "--write-splits") splits_dir = next();
    }
    try {
        if (!dict_dir.empty()) {
            // examples/matcher.rs:19-52: a dictionary of whole files, every query file matched against it unless it is
            // quieter than 0.03 (then silence); here the queries are matched in ONE batched call
            auto dictionary = SoundDictionary::from_path(dict_dir);
            dictionary->mode = dtw ? SS_DTW : SS_COSINE_REF;
            auto queries = SoundDictionary::from_path(query_dir);
            std::vector<std::shared_ptr<Sound>> loud;
            for (auto& q : queries->sounds)
                if (!(q->max_power() < 0.03)) loud.push_back(q);
            std::vector<uint32_t> idx;
            std::vector<double> dist;
            if (!loud.empty()) dictionary->match_indices(loud, {}, 1, &idx, &dist);
            printf("{\"dictionary\": %zu, \"queries\": %zu, \"silent\": %zu, \"matches\": [", dictionary->sounds.size(), queries->sounds.size(),
                   queries->sounds.size() - loud.size());
            for (size_t i = 0; i < loud.size(); i++)
                printf("%s[\"%s\", \"%s\"]", i ? "," : "", loud[i]->name.value_or("").c_str(), dictionary->sounds[idx[i]]->name.value_or("").c_str());
            printf("]}\n");
            return 0;
        }
        auto source = std::make_shared<Sound>(Sound::from_path(src));
        Partitioner partitioner(source);
        partitioner.set_threshold(threshold).set_depth(depth);
        if (!model_path.empty()) partitioner.train(GaussianMixtureModel::load(model_path));
        const std::vector<size_t> splits = partitioner.partition();  // "Must first train model" without -m
        if (!splits_dir.empty()) write_splits(*source, splits, splits_dir);  // examples/partition.rs:76
        if (partition_only) {
            printf("{\"source_frames\": %zu, \"max_power\": %.17g, \"nsplits\": %zu, \"splits\": [", source->num_frames(), source->max_power(), splits.size());
            for (size_t i = 0; i < splits.size(); i++) printf("%s%zu", i ? "," : "", splits[i]);
            printf("]}\n");
            return 0;
        }
        auto dictionary = SoundDictionary::from_segments(*source, splits);
        dictionary->mode = dtw ? SS_DTW : SS_COSINE_REF;

        auto target = std::make_shared<Sound>(Sound::from_path(tgt));
        partitioner.sound = target;
        const std::vector<size_t> tsplits = partitioner.partition();
        std::vector<std::shared_ptr<Sound>> segments;  // examples/reconstruction.rs:77-81
        size_t spos = 0, fpos = 0;
        for (size_t sp : tsplits) {
            std::vector<double> samp(target->samples().begin() + spos, target->samples().begin() + spos + sp);
            std::vector<double> m(target->mfccs().begin() + fpos, target->mfccs().begin() + fpos + sp / HOP * NCOEFFS);
            spos += sp;
            fpos += sp / HOP * NCOEFFS;
            segments.push_back(std::make_shared<Sound>(Sound::from_cut(std::move(samp), target->sample_rate(), std::move(m), target->context())));
        }
        SoundSequence sequence(segments);
        std::vector<uint32_t> idx;
        Sound result = sequence.clone_from_dictionary_to_sound(*dictionary, &idx);
        if (!out.empty()) result.write_file(out);
        double sum = 0;
        for (double s : result.samples()) sum += s;
        printf("{\"nsplits\": %zu, \"ntarget\": %zu, \"out_samples\": %zu, \"out_sum\": %.17g, \"out_frames\": %zu, \"idx\": [", splits.size(),
               tsplits.size(), result.samples().size(), sum, result.num_frames());
        for (size_t i = 0; i < idx.size(); i++) printf("%s%u", i ? "," : "", idx[i]);
        printf("]}\n");
    } catch (const CosError& e) {
        printf("{\"error\": \"%s\", \"code\": %d}\n", e.what(), e.code);
        return 1;
    }
    return 0;
}
