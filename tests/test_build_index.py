"""CPU suite: the corpus side of build_index (reference build_fm_index.py:37-114): TSV dialects, text clean-up, labels,
stand-in tokenizer.  The index build itself is GPU work and is covered by tests/test_gpu_parity.py."""
import json
import os

import pytest

from fm_index_constrained_decoding_b200 import build_index as bi

HERE = os.path.dirname(os.path.abspath(__file__))


def test_kilt_and_dpr_dialects(tmp_path):
    kilt = tmp_path / "k.tsv"
    kilt.write_text("7\tTitle One\tfirst   text\there  BULLET::::x\n"
                    "bad line without tabs\n"
                    "8\tEmpty\t   \n"
                    " 9 \t Second \tSECTION::::Hello World\n", encoding="utf-8")
    docs = list(bi.iter_documents(str(kilt)))
    assert docs == [("7", "first text here x"), ("9", "Hello World")]
    docs = list(bi.iter_documents(str(kilt), include_title=True, delim="@@", lowercase=True))
    assert docs == [("7", "title one @@ first text here x"), ("9", "second @@ hello world")]
    dpr = tmp_path / "d.tsv"
    dpr.write_text('id\ttext\ttitle\n1\t"some ""quoted"" text"\tT1\n2\tplain\tT2\n', encoding="utf-8")
    assert list(bi.iter_documents(str(dpr), fmt="dpr")) == [("1", 'some "quoted" text'), ("2", "plain")]
    with pytest.raises(ValueError):
        list(bi.iter_documents(str(kilt), fmt="other"))


def test_standin_tokenizer_and_cli_arguments():
    ids = bi.standin_encode("alpha beta alpha")
    assert len(ids) == 1 + 3 + 1 and ids[-1] == bi.STANDIN_EOS and ids[1] == ids[3] != ids[2]
    assert all(0 <= t < 128000 for t in ids[:-1])
    a = bi.parse_args(["in.tsv", "out/prefix", "--format", "dpr", "--include_title", "--delim", "##", "--lowercase",
                       "--jobs", "4", "--hf_model", "some/model"])
    assert (a.input, a.output, a.format, a.include_title, a.delim, a.lowercase, a.jobs, a.hf_model) == \
        ("in.tsv", "out/prefix", "dpr", True, "##", True, 4, "some/model")
    with pytest.raises(SystemExit):
        bi.main(["in.tsv", "out"])  # neither tokenizer given


@pytest.mark.skipif(not os.path.exists("/root/reference/test_data/sample_corpus.tsv"), reason="reference tree not present")
def test_sample_corpus_matches_the_golden_token_ids():
    """configs[0]: the documents of the reference's shipped sample corpus, tokenised by the stand-in, are exactly the
    token lists the golden vectors were generated from."""
    gold = json.load(open(os.path.join(HERE, "golden", "sample_corpus.json")))
    docs = list(bi.iter_documents("/root/reference/test_data/sample_corpus.tsv", literal_tabs=True))
    assert [d[0] for d in docs] == gold["labels"]
    assert [bi.standin_encode(d[1]) for d in docs] == gold["docs"]
