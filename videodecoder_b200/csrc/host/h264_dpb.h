// h264_dpb.h -- decoded picture buffer of the host side: which frame-pool slot of the reconstruction engine a picture is
// written to, which slots RefPicList0 names, when a slot leaves the reference set, in which order pictures are output and when a
// slot may be handed back (h264r_release).  Progressive frames, P / I slices: ITU-T H.264 8.2.1 (picture order count types 0 and
// 2), 8.2.4 (picture numbers, initialisation and modification of RefPicList0), 8.2.5 (sliding window and memory management
// control operations 1-6), C.4.5.3 (bumping, with a configurable reorder depth).
//
// Replaces, on the reference's side, Decoder (h264/Decoder.cpp:133-497): init_RefPicList / opra_RefModfication (short-term only;
// the long-term branch is an empty stub, h264/Decoder.cpp:324-326), calc_PictureNum, ctrl_FIFO, ctrl_MMOC_1..6, the POC-ordered
// output stack out_DecodedPic and ctrl_Memory; and Slice::Calc_POC (h264/slice.cpp:391-443, type 0 only).  The reference owns
// `picture` objects; here a picture is a slot number of the engine's frame pool (h264r_frame_begin's frame_id).
#pragma once
#include <stdint.h>
#include <vector>

#include "h264_headers.h"

namespace h264host {

struct DpbFrame {
    int slot = -1;
    bool is_ref = false, long_term = false, waiting_output = false;
    int frame_num = 0, frame_num_wrap = 0, pic_num = 0, long_term_idx = 0, long_term_pic_num = 0;
    int poc = 0;
    int decode_index = 0;
};

class Dpb {
public:
    // pool: number of frame slots (0 = never reuse: slot = index of the picture in decoding order); reorder: pictures held back
    // for output reordering (0 = output in decoding order as soon as decoded)
    void configure(int pool, int reorder) { pool_ = pool; reorder_ = reorder; }
    void reset()
    {
        frames_.clear();
        decoded_ = 0;
        prev_poc_msb_ = prev_poc_lsb_ = 0; prev_frame_num_offset_ = 0; prev_frame_num_ = 0; prev_had_mmco5_ = false;
        max_long_term_idx_ = -1;
    }
    const char* error() const { return err_; }

    // Before the picture's slice data: POC, slot, RefPicList0.  Returns 0 or a negative error code.
    int begin_picture(const Sps& sps, const SliceHeader& h, int* slot_out, int ref_list0[32], int* n_refs)
    {
        if (h.idr()) {
            // all reference pictures are marked unused; pictures still waiting for output are output first (no_output_of_prior_pics_flag
            // is not honoured: nothing is discarded)
            for (auto& f : frames_) f.is_ref = false;
            bump_all();
            prev_poc_msb_ = prev_poc_lsb_ = 0;
            prev_frame_num_offset_ = 0;
            max_long_term_idx_ = -1;
        }
        cur_ = DpbFrame();
        cur_.frame_num = h.frame_num;
        cur_.decode_index = decoded_;
        int rc = compute_poc(sps, h);
        if (rc) return rc;
        collect_garbage();
        // slot
        if (pool_ == 0) cur_.slot = decoded_;
        else {
            cur_.slot = -1;
            for (int s = 0; s < pool_ && cur_.slot < 0; s++) {
                bool busy = false;
                for (auto& f : frames_) busy |= f.slot == s;
                for (int r : pending_release_) busy |= r == s;
                if (!busy) cur_.slot = s;
            }
            if (cur_.slot < 0) {
                // a slot whose release is pending may be taken: the engine orders an overwrite behind queued readers
                if (!pending_release_.empty()) { cur_.slot = pending_release_.front(); pending_release_.erase(pending_release_.begin()); }
                else { err_ = "frame pool exhausted (pool smaller than max_num_ref_frames + reorder depth + 1)"; return kErrCapacity; }
            }
        }
        *slot_out = cur_.slot;
        *n_refs = 0;
        if (h.slice_type == kSliceP) {
            rc = build_list0(sps, h, ref_list0, n_refs);
            if (rc) return rc;
        }
        return kOk;
    }

    // After the picture: reference marking (8.2.5), output queue.  Fills the slots to output now (display order) and the slots that
    // are neither referenced nor waiting for output any more.
    int end_picture(const Sps& sps, const SliceHeader& h, std::vector<DpbFrame>* output, std::vector<int>* release)
    {
        bool mmco5 = false;
        if (h.nal_ref_idc != 0) {
            if (h.idr()) {
                cur_.is_ref = true;
                if (h.long_term_reference_flag) { cur_.long_term = true; cur_.long_term_idx = 0; max_long_term_idx_ = 0; }
                else max_long_term_idx_ = -1;
            } else {
                if (h.adaptive_marking) {
                    const int rc = run_mmco(sps, h, &mmco5);
                    if (rc) return rc;
                } else sliding_window(sps);
                cur_.is_ref = true;
            }
        }
        if (mmco5) {
            // 8.2.1: after memory_management_control_operation 5 the picture is treated as having frame_num 0 and POC tempPicOrderCnt = 0
            bump_all();
            cur_.frame_num = 0;
            cur_.poc = 0;
        }
        cur_.waiting_output = true;
        frames_.push_back(cur_);
        decoded_++;
        prev_frame_num_ = cur_.frame_num;
        prev_had_mmco5_ = mmco5;
        if (h.nal_ref_idc != 0) {
            if (mmco5) { prev_poc_msb_ = 0; prev_poc_lsb_ = 0; }
            else { prev_poc_msb_ = cur_poc_msb_; prev_poc_lsb_ = h.poc_lsb; }
        }
        // bumping
        for (;;) {
            int waiting = 0;
            for (auto& f : frames_) waiting += f.waiting_output;
            if (waiting <= reorder_) break;
            bump_one();
        }
        collect_garbage();
        hand_over(output, release);
        return kOk;
    }
    // end of stream: everything still waiting is output
    void flush(std::vector<DpbFrame>* output, std::vector<int>* release)
    {
        bump_all();
        for (auto& f : frames_) f.is_ref = false;
        collect_garbage();
        hand_over(output, release);
    }
    int current_poc() const { return cur_.poc; }
    int num_frames() const { return (int)frames_.size(); }

private:
    void hand_over(std::vector<DpbFrame>* output, std::vector<int>* release)
    {
        if (output) *output = out_;
        if (release) *release = pending_release_;
        out_.clear();
        pending_release_.clear();
    }
    int compute_poc(const Sps& sps, const SliceHeader& h)
    {
        if (sps.poc_type == 0) {
            const int max_lsb = 1 << sps.log2_max_poc_lsb;
            int msb;
            if (h.idr()) { prev_poc_msb_ = 0; prev_poc_lsb_ = 0; }
            if (h.poc_lsb < prev_poc_lsb_ && prev_poc_lsb_ - h.poc_lsb >= max_lsb / 2) msb = prev_poc_msb_ + max_lsb;
            else if (h.poc_lsb > prev_poc_lsb_ && h.poc_lsb - prev_poc_lsb_ > max_lsb / 2) msb = prev_poc_msb_ - max_lsb;
            else msb = prev_poc_msb_;
            cur_poc_msb_ = msb;
            const int top = msb + h.poc_lsb, bottom = top + h.delta_poc_bottom;
            cur_.poc = top < bottom ? top : bottom;
            return kOk;
        }
        if (sps.poc_type == 2) {
            const int max_frame_num = 1 << sps.log2_max_frame_num;
            int offset;
            if (h.idr()) offset = 0;
            else if (prev_had_mmco5_) offset = 0 + (prev_frame_num_ > h.frame_num ? max_frame_num : 0);
            else offset = prev_frame_num_offset_ + (prev_frame_num_ > h.frame_num ? max_frame_num : 0);
            prev_frame_num_offset_ = offset;
            cur_.poc = h.idr() ? 0 : (h.nal_ref_idc == 0 ? 2 * (offset + h.frame_num) - 1 : 2 * (offset + h.frame_num));
            return kOk;
        }
        err_ = "pic_order_cnt_type 1";
        return kErrUnsupported;
    }

    void picture_numbers(const Sps& sps, const SliceHeader& h)
    {
        const int max_frame_num = 1 << sps.log2_max_frame_num;
        for (auto& f : frames_) {
            if (!f.is_ref) continue;
            if (f.long_term) f.long_term_pic_num = f.long_term_idx;
            else {
                f.frame_num_wrap = f.frame_num > h.frame_num ? f.frame_num - max_frame_num : f.frame_num;
                f.pic_num = f.frame_num_wrap;
            }
        }
    }

    int build_list0(const Sps& sps, const SliceHeader& h, int list[32], int* n_refs)
    {
        picture_numbers(sps, h);
        std::vector<const DpbFrame*> l;
        // 8.2.4.2.1: short-term by descending PicNum, then long-term by ascending LongTermPicNum
        std::vector<const DpbFrame*> st, lt;
        for (auto& f : frames_) if (f.is_ref) (f.long_term ? lt : st).push_back(&f);
        for (size_t i = 1; i < st.size(); i++)
            for (size_t j = i; j > 0 && st[j - 1]->pic_num < st[j]->pic_num; j--) std::swap(st[j - 1], st[j]);
        for (size_t i = 1; i < lt.size(); i++)
            for (size_t j = i; j > 0 && lt[j - 1]->long_term_pic_num > lt[j]->long_term_pic_num; j--) std::swap(lt[j - 1], lt[j]);
        l = st;
        l.insert(l.end(), lt.begin(), lt.end());
        const int n = h.num_ref_idx_l0;
        l.resize((size_t)n + 1, nullptr);
        // 8.2.4.3: modification
        const int max_pic_num = 1 << sps.log2_max_frame_num, curr_pic_num = h.frame_num;
        int pred = curr_pic_num, idx = 0;
        for (int k = 0; k < h.n_mods && idx < n; k++) {
            const RefListMod& m = h.mods[k];
            const DpbFrame* target = nullptr;
            if (m.idc == 0 || m.idc == 1) {
                const int diff = m.value + 1;
                int no_wrap;
                if (m.idc == 0) no_wrap = pred - diff < 0 ? pred - diff + max_pic_num : pred - diff;
                else no_wrap = pred + diff >= max_pic_num ? pred + diff - max_pic_num : pred + diff;
                pred = no_wrap;
                const int pic_num = no_wrap > curr_pic_num ? no_wrap - max_pic_num : no_wrap;
                for (auto& f : frames_) if (f.is_ref && !f.long_term && f.pic_num == pic_num) target = &f;
            } else {
                for (auto& f : frames_) if (f.is_ref && f.long_term && f.long_term_pic_num == m.value) target = &f;
            }
            if (!target) { err_ = "ref_pic_list_modification names a picture that is not in the DPB"; return kErrBitstream; }
            for (int c = n; c > idx; c--) l[(size_t)c] = l[(size_t)c - 1];
            l[(size_t)idx++] = target;
            int w = idx;
            for (int c = idx; c <= n; c++)
                if (l[(size_t)c] != target) l[(size_t)w++] = l[(size_t)c];
            for (; w <= n; w++) l[(size_t)w] = nullptr;
        }
        *n_refs = n;
        for (int i = 0; i < 32; i++) list[i] = i < n && l[(size_t)i] ? l[(size_t)i]->slot : -1;
        if (n > 0 && list[0] < 0) { err_ = "P slice without a reference picture"; return kErrBitstream; }
        // entries beyond the available pictures repeat the first one: an index into them is a stream error, not a crash
        for (int i = 0; i < n; i++) if (list[i] < 0) list[i] = list[0];
        return kOk;
    }

    void sliding_window(const Sps& sps)
    {
        int n_short = 0, n_long = 0;
        for (auto& f : frames_) if (f.is_ref) (f.long_term ? n_long : n_short)++;
        const int max_refs = sps.max_num_ref_frames > 0 ? sps.max_num_ref_frames : 1;
        if (n_short + n_long >= max_refs && n_short > 0) {
            DpbFrame* oldest = nullptr;
            const int max_frame_num = 1 << sps.log2_max_frame_num;
            for (auto& f : frames_) {
                if (!f.is_ref || f.long_term) continue;
                f.frame_num_wrap = f.frame_num > cur_.frame_num ? f.frame_num - max_frame_num : f.frame_num;
                if (!oldest || f.frame_num_wrap < oldest->frame_num_wrap) oldest = &f;
            }
            oldest->is_ref = false;
        }
    }

    int run_mmco(const Sps& sps, const SliceHeader& h, bool* mmco5)
    {
        picture_numbers(sps, h);
        const int curr_pic_num = h.frame_num;
        for (int i = 0; i < h.n_mmco; i++) {
            const Mmco& m = h.mmco[i];
            switch (m.op) {
            case 1: {
                const int pic_num = curr_pic_num - (m.a + 1);
                for (auto& f : frames_) if (f.is_ref && !f.long_term && f.pic_num == pic_num) f.is_ref = false;
                break;
            }
            case 2:
                for (auto& f : frames_) if (f.is_ref && f.long_term && f.long_term_pic_num == m.a) f.is_ref = false;
                break;
            case 3: {
                const int pic_num = curr_pic_num - (m.a + 1);
                for (auto& f : frames_) if (f.is_ref && f.long_term && f.long_term_idx == m.b) f.is_ref = false;
                for (auto& f : frames_)
                    if (f.is_ref && !f.long_term && f.pic_num == pic_num) { f.long_term = true; f.long_term_idx = m.b; f.long_term_pic_num = m.b; }
                break;
            }
            case 4:
                max_long_term_idx_ = m.a - 1;
                for (auto& f : frames_) if (f.is_ref && f.long_term && f.long_term_idx > max_long_term_idx_) f.is_ref = false;
                break;
            case 5:
                for (auto& f : frames_) f.is_ref = false;
                max_long_term_idx_ = -1;
                *mmco5 = true;
                break;
            case 6:
                for (auto& f : frames_) if (f.is_ref && f.long_term && f.long_term_idx == m.b) f.is_ref = false;
                cur_.long_term = true;
                cur_.long_term_idx = m.b;
                break;
            default: break;
            }
        }
        (void)sps;
        return kOk;
    }

    void bump_one()
    {
        DpbFrame* best = nullptr;
        for (auto& f : frames_)
            if (f.waiting_output && (!best || f.poc < best->poc)) best = &f;
        if (!best) return;
        best->waiting_output = false;
        out_.push_back(*best);
    }
    void bump_all()
    {
        for (;;) {
            bool any = false;
            for (auto& f : frames_) any |= f.waiting_output;
            if (!any) break;
            bump_one();
        }
    }
    void collect_garbage()
    {
        for (size_t i = 0; i < frames_.size();) {
            if (!frames_[i].is_ref && !frames_[i].waiting_output) {
                if (pool_ != 0) pending_release_.push_back(frames_[i].slot);
                frames_.erase(frames_.begin() + (long)i);
            } else i++;
        }
    }

    int pool_ = 0, reorder_ = 0;
    std::vector<DpbFrame> frames_;
    std::vector<int> pending_release_;
    std::vector<DpbFrame> out_;
    DpbFrame cur_;
    int decoded_ = 0;
    int prev_poc_msb_ = 0, prev_poc_lsb_ = 0, cur_poc_msb_ = 0, prev_frame_num_offset_ = 0, prev_frame_num_ = 0;
    bool prev_had_mmco5_ = false;
    int max_long_term_idx_ = -1;
    const char* err_ = "";
};

}  // namespace h264host
