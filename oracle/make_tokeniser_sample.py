"""Adds the reference's ``tokens_lemmatized`` column to tests/golden/tokeniser_sample.json.

TEST INFRASTRUCTURE (build container only: reads /root/reference/data/test_queries_bm25.csv).
The 60 sample queries were frozen in round 1 with their ``tokens_regular``; the lemmatised tokens
of the same rows pin the offline lemma path (a-nice-rag_b200/processing/preprocess_bm25.py).
"""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SAMPLE = os.path.join(ROOT, "tests", "golden", "tokeniser_sample.json")

if __name__ == "__main__":
    with open("/root/reference/data/test_queries_bm25.csv", encoding="utf-8") as fh:
        by_query = {}
        for row in csv.DictReader(fh):
            by_query.setdefault((row["query"], row["tokens_regular"]), row["tokens_lemmatized"])
    with open(SAMPLE) as fh:
        sample = json.load(fh)
    for row in sample["rows"]:
        row["tokens_lemmatized"] = by_query[(row["query"], row["tokens_regular"])]
    if "make_tokeniser_sample" not in sample["source"]:
        sample["source"] += "; tokens_lemmatized added by oracle/make_tokeniser_sample.py"
    with open(SAMPLE, "w") as fh:
        json.dump(sample, fh, indent=1)
        fh.write("\n")
    print(f"{SAMPLE}: {len(sample['rows'])} rows")
