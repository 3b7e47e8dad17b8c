"""Writes tests/golden/kat1_fig3.json: the worked example of the GCSA2 paper.

Source: reference paper/gcsa2_graph_dbg.ipe:305-316, 453-488 (input graph, Figure 2) and
paper/gcsa2_pruned_index.ipe (Figure 3: order-3 pruned de Bruijn graph and its GCSA),
transcribed in SURVEY.md section 4.  Nothing is computed here: the arrays are the figure's,
the expected answers are the ones the figure and its caption (paper/paper.tex:305) state.
Values "0:1" / "0:2" of the figure are the source node's extra positions; locate() derives
an unsampled node's value as sample + steps (src/gcsa.cpp:893), so "0:2" + 1 = "0:1" and
"0:1" + 1 = 0, i.e. they are -1 and -2 in 64-bit arithmetic.
"""
import json, os

M1, M2 = (1 << 64) - 1, (1 << 64) - 2
kat = {
    "order": 3,
    "keys":   ["$$$", "A$", "ATA", "ATC", "ATG", "CA", "CT", "GC", "GT", "TA", "TC", "TG", "TT", "#G", "##G", "###"],
    "values": [[11], [10], [7], [3], [3], [2, 6], [2], [1], [8], [9], [5], [5], [4], [0], [M1], [M2]],
    # comp order $ A C G T N #  (src/support.cpp:92)
    "bwt": {"0": [15], "1": [0, 9, 10, 11], "2": [2, 3, 4, 12], "3": [5, 6, 9], "4": [1, 5, 8, 10, 11], "5": [], "6": [7, 13, 14]},
    "C": [0, 1, 5, 9, 12, 17, 17, 20],
    "edges": "11111001101111101111",
    "sampled_paths": [2, 3, 4, 5, 8, 9, 10, 11, 12, 15],
    "stored_samples": [7, 3, 3, 2, 6, 8, 9, 5, 5, 4, M2],
    "samples": "11101111111",
    "graph": {"labels": "#GCATTCAGTA$", "edges": [[0, 1], [1, 2], [2, 3], [2, 4], [3, 5], [4, 5], [5, 6], [5, 8],
                                                 [6, 7], [7, 9], [8, 9], [9, 10], [10, 11]]},
    "find": {"AT": [2, 4], "CAT": [5, 5], "CA": [5, 5], "T": [9, 12], "TA": [9, 9], "$#": [0, 0], "GCA": [7, 7],
             "GCT": [7, 7], "ATT": None, "$$": None, "": [0, 15]},
    "lf": [{"range": [9, 12], "char": "A", "result": [2, 4]}, {"range": [9, 9], "char": "G", "result": [8, 8]}],
    "locate": {"CA": [2, 6], "AT": [3, 7], "GC": [1], "A": [3, 7, 10], "#G": [0]},
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat1_fig3.json"), "w") as f:
    json.dump(kat, f, indent=1)
