"""Corpus -> index file, with the command line of the reference's build_fm_index.py (:117-139):

    python -m fm_index_constrained_decoding_b200.build_index INPUT.tsv OUTPUT_PREFIX \
        [--format kilt|dpr] [--include_title] [--delim @@] [--lowercase] [--hf_model NAME | --standin_tokenizer]

What it keeps from the reference (build_fm_index.py:37-114): the two TSV dialects (kilt: id<TAB>title<TAB>text per line;
dpr: header line, csv with id, text, title), whitespace runs collapsed, `BULLET::::` / `SECTION::::` removed, empty texts
skipped, `title delim text` when --include_title, labels = the ids in file order, and per document the token ids of
"Paraphrase: " + text followed by EOS (:46-48; a leading blank for BART models).  What differs: the suffix sort, BWT and
wavelet matrix are built on the GPU (seconds instead of an sdsl construct), `ftfy.fix_text` is applied only when ftfy is
importable, `--tokenize` needs spaCy like the reference, and when no Hugging Face tokenizer is available offline
`--standin_tokenizer` gives the documented stand-in (whitespace split, FNV-1a word hash mod 128,000, EOS 128009) that the
golden vectors of tests/golden/sample_corpus.json were made with.
"""
import argparse
import csv
import re

STANDIN_EOS = 128009


def _fix_text(text):
    try:
        import ftfy
    except ImportError:
        return text
    return ftfy.fix_text(text)


def iter_documents(path, fmt="kilt", include_title=False, delim="@@", lowercase=False, word_tokenize=None,
                   literal_tabs=False):
    """Yields (label, text) like build_fm_index.py:53-97.  literal_tabs: treat the two characters backslash-t as TAB
    (the reference's shipped test_data/sample_corpus.tsv is written that way)."""
    with open(path, "r", 2 ** 16, encoding="utf-8") as f:
        if fmt == "dpr":
            next(f)
            rows = ((pp[0], pp[2], pp[1]) for pp in csv.reader(f, delimiter="\t", quotechar='"') if len(pp) == 3)
        elif fmt == "kilt":
            lines = (line.strip() for line in f)
            if literal_tabs:
                lines = (line.replace("\\t", "\t") for line in lines)
            rows = (tuple(pp) for pp in (line.split("\t", 2) for line in lines) if len(pp) == 3)
        else:
            raise ValueError("format must be kilt or dpr")
        for idx, title, text in rows:
            idx, title = idx.strip(), title.strip()
            text = _fix_text(re.sub(r"\s+", " ", text)).replace("BULLET::::", "").replace("SECTION::::", "").strip()
            if not text:
                continue
            if word_tokenize is not None:
                title, text = " ".join(word_tokenize(title)), " ".join(word_tokenize(text))
            title = "%s %s" % (title, delim)
            if include_title and title:
                text = "%s %s" % (title, text)
            if lowercase:
                text = text.lower()
            yield idx, text


def standin_encode(text):
    """The stand-in tokenizer of tests/golden/make_sample_corpus_golden.py: one id per whitespace-separated word."""
    out = []
    for word in ("Paraphrase: " + text).split():
        h = 0xCBF29CE484222325
        for b in word.encode("utf-8"):
            h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
        out.append(h % 128000)
    return out + [STANDIN_EOS]


def hf_encoder(model_name):
    """build_fm_index.py:29-48 for a Hugging Face tokenizer (local files only: there is no network here)."""
    from transformers import AutoTokenizer
    tok = AutoTokenizer.from_pretrained(model_name, use_fast=False, local_files_only=True)
    is_bart = "bart" in model_name

    def encode(text):
        line = text.strip()
        if is_bart:
            line = " " + line
        return tok("Paraphrase: " + line, add_special_tokens=False)["input_ids"] + [tok.eos_token_id]

    return encode


def build_index(input_path, encode, fmt="kilt", include_title=False, delim="@@", lowercase=False, word_tokenize=None,
                literal_tabs=False, layout=None):
    """build_fm_index.py:100-114: returns an initialised FMIndex with .labels set."""
    from .index import FMIndex
    labels, docs = [], []
    for label, text in iter_documents(input_path, fmt, include_title, delim, lowercase, word_tokenize, literal_tabs):
        labels.append(label)
        docs.append(encode(text))
    index = FMIndex(layout=layout) if layout else FMIndex()
    index.initialize(docs, in_memory=True)
    index.labels = labels
    return index


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("input")
    ap.add_argument("output")
    ap.add_argument("--jobs", type=int, default=1, help="accepted for compatibility; tokenisation runs in this process")
    ap.add_argument("--include_title", action="store_true")
    ap.add_argument("--delim", default="@@")
    ap.add_argument("--format", choices=["kilt", "dpr"], default="kilt")
    ap.add_argument("--hf_model", default=None, type=str)
    ap.add_argument("--lowercase", action="store_true")
    ap.add_argument("--tokenize", action="store_true")
    ap.add_argument("--standin_tokenizer", action="store_true")
    ap.add_argument("--literal_tabs", action="store_true")
    ap.add_argument("--layout", default=None, choices=["resident", "sampled"])
    return ap.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    word_tokenize = None
    if args.tokenize:
        from spacy.lang.en import English
        nlp = English().tokenizer

        def word_tokenize(text):
            return [t.text.strip() for t in nlp(text)]
    if args.standin_tokenizer:
        encode = standin_encode
    elif args.hf_model is not None:
        encode = hf_encoder(args.hf_model)
    else:
        raise SystemExit("give --hf_model NAME (tokenizer files must be available locally) or --standin_tokenizer; the "
                         "reference's default (fairseq BART from torch.hub) needs the network")
    index = build_index(args.input, encode, args.format, args.include_title, args.delim, args.lowercase, word_tokenize,
                        args.literal_tabs, args.layout)
    index.save(args.output)
    print("indexed %d documents, %d tokens -> %s.fmi / .oth" % (len(index.labels), len(index), args.output))


if __name__ == "__main__":
    main()
