"""Generate tests/golden/cfg1_collator.pt: BASELINE.json configs[0] end to end through the REFERENCE's own data path
(test infrastructure; runs only in the build container).

A synthetic OK-VQA reader JSON in the producer's schema (data_process/vqa2_deal.py:137-144: question, target, answer,
img_id, caption, fact[{sentence, id}]) built from the questions / captions the reference ships, is pushed through the
reference's unmodified `src.data.Dataset` + `src.data.Collator` (src/data.py:14-148) with a stand-in tokenizer (no T5
vocabulary exists offline: words hash into [11, 32000), ':' -> 10, '.' -> 5, '</s>' -> 1, pad 0 -- the ids
src/model.py:102,127,178 relies on) at n_context=2, text_maxlength=64, answer_maxlength=10, stream=2, and the
resulting tensors go through the reference `FiDT5` (t5-base shape, seeded random init) exactly as test_reader.py:60-76
drives it: forward loss, greedy generate(max_length=10) and get_crossattention_scores.

    python oracle/gen_collator_golden.py
"""
import json
import os
import re
import sys
import types
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


class StandInTokenizer:
    """Whitespace/punctuation tokenizer with the id conventions the reader path depends on."""
    pad_token_id, eos_token_id = 0, 1

    def tokens(self, text):
        text = text.replace("</s>", " \x01 ")
        return re.findall(r"\x01|\w+|[^\w\s]", text)

    def encode(self, text):
        out = []
        for t in self.tokens(text):
            if t == "\x01":
                out.append(1)
            elif t == ":":
                out.append(10)
            elif t == ".":
                out.append(5)
            else:
                out.append(11 + zlib.crc32(t.lower().encode()) % 31989)
        return out

    def batch_encode_plus(self, texts, max_length=None, pad_to_max_length=False, return_tensors=None, truncation=False):
        enc = [self.encode(t) for t in texts]
        if truncation and max_length is not None:
            enc = [e[:max_length] for e in enc]
        width = max_length if (max_length is not None and pad_to_max_length) else max(len(e) for e in enc)
        ids = torch.zeros(len(enc), width, dtype=torch.long)
        mask = torch.zeros(len(enc), width, dtype=torch.long)
        for i, e in enumerate(enc):
            ids[i, :len(e)] = torch.tensor(e, dtype=torch.long)
            mask[i, :len(e)] = 1
        return {"input_ids": ids, "attention_mask": mask}


def synthetic_reader_json(n=3):
    """Reader examples in the schema of data_process/vqa2_deal.py:137-144 from the shipped questions and captions."""
    base = "/root/reference/data_process/data/okvqa"
    with open(os.path.join(base, "cache/1/valid.json")) as f:
        valid = json.load(f)
    with open("/root/reference/data_process/data/caption/pred.coco_caption.val.json") as f:
        caps = json.load(f)
    cap_by_img = {}
    for c in (caps if isinstance(caps, list) else caps.get("annotations", caps)):
        if isinstance(c, dict) and "image_id" in c:
            cap_by_img[int(c["image_id"])] = c.get("caption", "")
    items = valid if isinstance(valid, list) else list(valid.values())
    out = []
    for it in items:
        img = it.get("img_id", it.get("image_id"))
        try:
            img_num = int(re.findall(r"\d+", str(img))[-1])
        except Exception:
            continue
        cap = cap_by_img.get(img_num)
        q = it.get("sent", it.get("question"))
        lab = it.get("label", it.get("answer"))
        if not cap or not q or not lab:
            continue
        answers = lab if isinstance(lab, dict) else {str(lab): 1.0}
        target = max(answers, key=answers.get)
        words = [w for w in re.findall(r"\w+", q.lower()) if len(w) > 3] or ["thing"]
        facts = [{"sentence": f"{w} is related to {target} .", "id": f"syn{j}"} for j, w in enumerate(words[:4])]
        facts.append({"sentence": f"a {target} is used for something", "id": "syn_tail"})  # no final '.': unterminated fact
        out.append({"question": q, "target": target, "answer": answers, "img_id": str(img), "caption": cap, "fact": facts})
        if len(out) == n:
            break
    assert len(out) == n, "could not assemble the synthetic reader JSON from the shipped OK-VQA files"
    return out


def collate_with_reference(examples, n_context=2, text_maxlength=64, answer_maxlength=10):
    import src.data as ref_data
    opt = types.SimpleNamespace(n_context=n_context, fact_use_way="concate", use_fact="yes")
    ds = ref_data.Dataset(examples, opt)
    col = ref_data.Collator(text_maxlength, StandInTokenizer(), answer_maxlength=answer_maxlength, stream=2)
    return col([ds[i] for i in range(len(ds))])


def main():
    from oracle import fidt5_oracle as O
    from oracle.gen_golden import build_reference, install_capture_hooks
    examples = synthetic_reader_json(3)
    index, target_ids, target_mask, passage_ids, passage_masks = collate_with_reference(examples)
    cfg = O.T5_BASE
    seed, emb_std = 300, 0.25
    sd = O.perturb_norm_weights(O.init_state_dict(cfg, seed=seed, emb_std=emb_std), seed=seed + 1)
    model = build_reference(cfg, sd)
    res = {"examples": examples, "seed": seed, "emb_std": emb_std, "n_context": 2, "text_maxlength": 64,
           "answer_maxlength": 10, "index": index, "target_ids": target_ids, "target_mask": target_mask,
           "passage_ids": passage_ids, "passage_masks": passage_masks, "per_example": []}
    install_capture_hooks(model)
    opt = types.SimpleNamespace(stream=2, n_context=2, attention_score_style="21mean", use_last_half_layer_attention="yes")
    for k in range(len(examples)):   # configs[0]: batch 1
        ids, mask, labels = passage_ids[k:k + 1], passage_masks[k:k + 1], target_ids[k:k + 1]
        with torch.no_grad():
            out = model(input_ids=ids, attention_mask=mask, labels=labels, return_dict=False)
            model.reset_score_storage()
            gen = model.generate(input_ids=ids, attention_mask=mask, max_length=10)
            sc = model.get_crossattention_scores(opt, ids, None, mask)
        ent = {"loss": out[0].detach(), "generated": gen, "scores_21mean_last_half": sc}
        if k == 0:
            ent["logits"] = out[1].detach()
        res["per_example"].append(ent)
    path = os.path.join(ROOT, "tests", "golden", "collator", "cfg1_collator.pt")
    torch.save(res, path)
    print(path, os.path.getsize(path), "bytes")
    for e, ent in zip(examples, res["per_example"]):
        print(e["question"], "|", float(ent["loss"]), ent["generated"].tolist(), ent["scores_21mean_last_half"].tolist())


if __name__ == "__main__":
    main()
