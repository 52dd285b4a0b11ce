# usage: ncu_txt.sh <name> <kernel regex> <rows per launch> <command...>   -- one full capture, exported as text, report deleted
name=$1; kre=$2; nrows=$3; shift 3
out=gpurun_out/$OUT
(timeout 300 ncu --set full --clock-control none -k regex:$kre -s 3 -c 1 -o $out/$name "$@") > $out/ncu_$name.log 2>&1
ncu -i $out/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>/dev/null
ncu -i $out/$name.ncu-rep --page source --csv > $out/${name}_sass.csv 2>/dev/null
python profiles/analyze_ncu.py $out/${name}_raw.csv $out/${name}_sass.csv $nrows > $out/ncu_$name.txt 2>&1
rm -f $out/$name.ncu-rep $out/${name}_sass.csv
head -c 20000000 $out/${name}_raw.csv > $out/${name}_raw.tmp && mv $out/${name}_raw.tmp $out/${name}_raw.csv
head -40 $out/ncu_$name.txt
