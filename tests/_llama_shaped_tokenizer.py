"""Builds a Llama-2-SHAPED fast tokenizer with the `tokenizers` library (the real vocabulary is gated and absent): 32000 entries,
`<unk>/<s>/</s>` = 0/1/2, the SentencePiece dummy prefix at id 29871, 256 action code points on ids 31744..31999, Metaspace with a prepended
prefix, BOS added by the post-processor. Shared by tests/test_real_tokenizer_path.py and oracle/gen_golden_processor_call.py."""
import json
import os


def build_llama_shaped_tokenizer(path: str) -> None:
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, processors

    vocab = {"<unk>": 0, "<s>": 1, "</s>": 2}
    text_chars = list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789:;,._-?!\n")
    for i, ch in enumerate(text_chars):
        vocab[ch] = 3 + i
    vocab["▁"] = 29871
    action_chars = [chr(0x4E00 + i) for i in range(256)]  # 256 distinct single code points (CJK block), like Llama-2's tail
    for i, ch in enumerate(action_chars):
        vocab[ch] = 31744 + i
    used = set(vocab.values())
    for i in range(32000):  # fill the remaining ids so that vocab_size == 32000
        if i not in used:
            vocab[f"<filler_{i}>"] = i
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="<unk>"))
    # "▁" is prepended to the text and every character is its own token (a WordLevel stand-in for SentencePiece pieces)
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Metaspace(replacement="▁", prepend_scheme="first", split=False),
                                                 pre_tokenizers.Split("", behavior="isolated")])  # fmt: skip
    tok.post_processor = processors.TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 1)])
    tok.decoder = decoders.Sequence([decoders.Replace("▁", " "), decoders.Fuse(), decoders.Strip(" ", 1, 0)])
    os.makedirs(path, exist_ok=True)
    tok.save(os.path.join(path, "tokenizer.json"))
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump({"tokenizer_class": "PreTrainedTokenizerFast", "bos_token": "<s>", "eos_token": "</s>", "unk_token": "<unk>",
                   "model_max_length": 2048, "clean_up_tokenization_spaces": False}, f)  # fmt: skip


